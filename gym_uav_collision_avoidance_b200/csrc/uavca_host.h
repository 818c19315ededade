// uavca_host.h — host-side declarations shared by the kernel translation unit and the C-ABI.
#pragma once

#include <cuda_runtime.h>

#include "uavca_device.cuh"

namespace uavca {

cudaError_t launch_step_multi(const KernelArgs& a, cudaStream_t st);
cudaError_t launch_reset_multi(const KernelArgs& a, const uint8_t* mask, cudaStream_t st);
cudaError_t launch_observe_multi(const KernelArgs& a, cudaStream_t st);
cudaError_t launch_step_single(const KernelArgs& a, cudaStream_t st);
cudaError_t launch_reset_single(const KernelArgs& a, const uint8_t* mask, cudaStream_t st);
cudaError_t launch_observe_single(const KernelArgs& a, cudaStream_t st);
cudaError_t launch_map_action(const Consts& c, const float* in, float* out, long long M, int mode, cudaStream_t st);
cudaError_t launch_stats(const StateView& s, int B, long long* out8, cudaStream_t st);

}  // namespace uavca
