// uavca_tma.cuh — the bulk path of the multi-UAV step: persistent CTAs, TMA bulk copies, mbarrier pipeline.
//
// A batch is cut into TILES of T = W * lanes UAVs (W compute warps, each holding floor(32/N) whole envs).  One CTA
// per SM slot walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... through an S-stage shared-memory ring:
//
//   IO warp (one elected lane)      full[s]            compute warps (one lane per UAV)        done[s]
//   -----------------------------   ----------------   -------------------------------------   ------------------
//   cp.async.bulk global -> stage   complete_tx  --->  wait, LDS inputs, step_core(),
//   (8 SoA arrays of the tile)                         STS outputs IN PLACE (pos/vel/prev/
//                                                      flags/steps) + obs/reward/done rows,
//                                                      fence.proxy.async, arrive        --->   wait,
//   cp.async.bulk stage -> global  <---------------------------------------------------------  8-9 bulk stores,
//   wait_group.read, refill stage                                                              commit_group
//
// Every byte of state and I/O crosses HBM <-> SM exactly once as 128 B..10 KB bulk transactions issued by one
// thread; the compute lanes execute no global load/store, no address arithmetic and no bounds predicate on the
// hot path, and memory latency is hidden by the stage ring instead of by occupancy.  Ragged tails (B not a
// multiple of the tile), odd lane counts and unaligned caller tensors go through the per-lane kernel in
// uavca_kernels.cu, which runs the very same step_core().
#pragma once

#include "uavca_multi.cuh"

namespace uavca {

#ifndef UAVCA_TMA_STAGES
#define UAVCA_TMA_STAGES 3
#endif
#ifndef UAVCA_TMA_MINB
#define UAVCA_TMA_MINB 4
#endif

template <int NT, bool FINAL>
struct TmaGeom {
  static constexpr int N = NT;
  static constexpr int EPW = 32 / N;        // envs per warp
  static constexpr int LANES = EPW * N;     // lanes of a warp that hold a UAV
  static constexpr bool kSupported = (LANES % 2) == 0;
  static constexpr int W = (LANES % 4 == 0) ? 4 : 8;  // compute warps: makes every tile slice a multiple of 16 bytes
  static constexpr int THREADS = (W + 1) * 32;
  static constexpr int MINB = W == 4 ? UAVCA_TMA_MINB : (UAVCA_TMA_MINB + 1) / 2;  // resident CTAs per SM aimed at
  static constexpr int T = W * LANES;       // UAVs per tile
  static constexpr int E = W * EPW;         // envs per tile
  static constexpr int S = UAVCA_TMA_STAGES;
  // byte offsets of the arrays inside one stage
  static constexpr int POS = 0;
  static constexpr int VEL = POS + T * 8;
  static constexpr int TGT = VEL + T * 16;
  static constexpr int ACT = TGT + T * 8;
  static constexpr int INIT = ACT + T * 8;
  static constexpr int PREV = INIT + T * 4;
  static constexpr int REW = PREV + T * 4;
  static constexpr int OBS = REW + T * 4;
  static constexpr int FOBS = OBS + T * 40;
  static constexpr int FLG = FOBS + (FINAL ? T * 40 : 0);
  static constexpr int DONE = FLG + T;
  static constexpr int STEPS = DONE + T;
  static constexpr int STAGE_BYTES = (STEPS + E * 4 + 127) / 128 * 128;
  static constexpr unsigned BYTES_IN = T * (8 + 16 + 8 + 8 + 4 + 4 + 1) + E * 4;
  static constexpr int RING_FLOATS = kRingFloats;  // per compute warp: position rings, candidates, headings
  static constexpr int SMEM_BYTES = S * STAGE_BYTES + W * RING_FLOATS * 4 + 2 * S * 8;
};

// ---- PTX: mbarrier, bulk copies, proxy fences ---------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "UAVCA_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra UAVCA_DONE_%=;\n"
      "bra UAVCA_WAIT_%=;\n"
      "UAVCA_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA store)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- step_core I/O policy over one shared-memory stage -----------------------------------------------------------

template <class G, bool FINAL>
struct SmemIO {
  const KernelArgs& a;
  unsigned char* st;  // stage base
  int u;              // UAV slot inside the tile
  int e;              // env slot inside the tile
  int m;              // flat UAV index in the shard (rare direct global stores)
  int env;            // env index in the shard
  bool valid;
  template <typename T>
  __device__ __forceinline__ T* at(int off) const { return reinterpret_cast<T*>(st + off); }
  __device__ __forceinline__ Uav load_uav() const {
    Uav r;
    const float2 p = at<float2>(G::POS)[u], t = at<float2>(G::TGT)[u];
    const double2 v = at<double2>(G::VEL)[u];
    r.px = p.x; r.py = p.y; r.tx = t.x; r.ty = t.y; r.vx = v.x; r.vy = v.y;
    r.init = at<float>(G::INIT)[u];
    r.prev = at<float>(G::PREV)[u];
    r.flags = at<uint8_t>(G::FLG)[u];
    return r;
  }
  __device__ __forceinline__ float2 load_action() const { return at<float2>(G::ACT)[u]; }
  __device__ __forceinline__ int load_steps() const { return at<int>(G::STEPS)[e]; }
  __device__ __forceinline__ void loads_done() const {}
  __device__ __forceinline__ void store_reward_done(float r, bool done) const {
    if (valid) {
      at<float>(G::REW)[u] = r;
      at<uint8_t>(G::DONE)[u] = (uint8_t)done;
    }
  }
  __device__ __forceinline__ bool wants_final() const { return FINAL; }
  bool final_open = FINAL;  // rows also go to the final_obs stage until commit_final()
  __device__ __forceinline__ void put_own(float2 o01, float2 o23) const {
    if (valid) {
      float2* row = reinterpret_cast<float2*>(st + G::OBS + u * 40);
      row[0] = o01; row[1] = o23;
      if (FINAL && final_open) {
        float2* frow = reinterpret_cast<float2*>(st + G::FOBS + u * 40);
        frow[0] = o01; frow[1] = o23;
      }
    }
  }
  __device__ __forceinline__ void put_neighbours(const ObsTail& n) const {
    if (valid) {
      float2* row = reinterpret_cast<float2*>(st + G::OBS + u * 40);
      row[2] = n.a; row[3] = n.b; row[4] = n.c;
      if (FINAL && final_open) {
        float2* frow = reinterpret_cast<float2*>(st + G::FOBS + u * 40);
        frow[2] = n.a; frow[3] = n.b; frow[4] = n.c;
      }
    }
  }
  __device__ __forceinline__ void commit_obs() const {}
  __device__ __forceinline__ void commit_final() { final_open = false; }
  __device__ __forceinline__ void store_state(const Uav& s) const {  // in place: the stage is drained by TMA stores
    if (valid) {
      at<float2>(G::POS)[u] = make_float2(s.px, s.py);
      at<double2>(G::VEL)[u] = make_double2(s.vx, s.vy);
      at<float>(G::PREV)[u] = s.prev;
      at<uint8_t>(G::FLG)[u] = (uint8_t)s.flags;
    }
  }
  __device__ __forceinline__ void store_target(const Uav& s) const {  // reset lanes only: straight to global
    a.s.tgt[m] = make_float2(s.tx, s.ty);
    a.s.init[m] = s.init;
  }
  __device__ __forceinline__ void store_steps(int v, bool leader) const {
    if (leader) at<int>(G::STEPS)[e] = v;
  }
  __device__ __forceinline__ void store_reset(bool rs) const {  // straight to global: one byte per env
    if (a.io.reset_mask) a.io.reset_mask[env] = (uint8_t)rs;
  }
};

// ---- the kernel ------------------------------------------------------------------------------------------------------

template <int NT, bool FINAL>
__global__ void __launch_bounds__(TmaGeom<NT, FINAL>::THREADS, TmaGeom<NT, FINAL>::MINB)
    step_multi_tma_kernel(const __grid_constant__ KernelArgs a, const int num_tiles) {
  using G = TmaGeom<NT, FINAL>;
  constexpr int S = G::S;
  extern __shared__ __align__(128) unsigned char dsm[];
  unsigned char* const stages = dsm;
  float* const rings = reinterpret_cast<float*>(dsm + S * G::STAGE_BYTES);
  const uint32_t bar0 = smem_u32(dsm + S * G::STAGE_BYTES + G::W * G::RING_FLOATS * 4);  // full[S], then done[S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(bar0 + 8 * s, 1);            // full: one arrive.expect_tx by the IO lane + the bytes
      mbar_init(bar0 + 8 * (S + s), G::W);   // done: one arrive per compute warp
    }
    fence_mbar_init();
  }
  __syncthreads();
  // Programmatic dependent launch: everything above overlapped with the tail of the previous kernel in the stream;
  // nothing below may (it reads what that kernel wrote).  Dependents may start launching right away: they wait in
  // their own cudaGridDependencySynchronize() until this grid has completed and flushed.
  cudaGridDependencySynchronize();
  cudaTriggerProgrammaticLaunchCompletion();

  const int n_my = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

  if (warp == G::W) {
    // ================= IO warp: one lane drives every bulk copy of this CTA =================
    if (lane == 0) {
      const uint32_t st0 = smem_u32(stages);
      auto issue_load = [&](int j) {
        const int s = j % S;
        const long long tile = (long long)blockIdx.x + (long long)j * gridDim.x;
        const long long m0 = tile * G::T, e0 = tile * G::E;
        const uint32_t sb = st0 + s * G::STAGE_BYTES, fb = bar0 + 8 * s;
        mbar_expect_tx(fb, G::BYTES_IN);
        bulk_g2s(sb + G::VEL, a.s.vel + m0, G::T * 16, fb);
        bulk_g2s(sb + G::POS, a.s.pos + m0, G::T * 8, fb);
        bulk_g2s(sb + G::TGT, a.s.tgt + m0, G::T * 8, fb);
        bulk_g2s(sb + G::ACT, a.io.action + m0, G::T * 8, fb);
        bulk_g2s(sb + G::INIT, a.s.init + m0, G::T * 4, fb);
        bulk_g2s(sb + G::PREV, a.s.prev + m0, G::T * 4, fb);
        bulk_g2s(sb + G::FLG, a.s.flags + m0, G::T, fb);
        bulk_g2s(sb + G::STEPS, a.s.steps + e0, G::E * 4, fb);
      };
      auto issue_store = [&](int j) {
        const int s = j % S;
        const long long tile = (long long)blockIdx.x + (long long)j * gridDim.x;
        const long long m0 = tile * G::T, e0 = tile * G::E;
        const uint32_t sb = st0 + s * G::STAGE_BYTES;
        bulk_s2g(a.io.obs + m0 * 10, sb + G::OBS, G::T * 40);
        if (FINAL) bulk_s2g(a.io.final_obs + m0 * 10, sb + G::FOBS, G::T * 40);
        bulk_s2g(a.s.vel + m0, sb + G::VEL, G::T * 16);
        bulk_s2g(a.s.pos + m0, sb + G::POS, G::T * 8);
        bulk_s2g(a.s.prev + m0, sb + G::PREV, G::T * 4);
        bulk_s2g(a.io.reward + m0, sb + G::REW, G::T * 4);
        bulk_s2g(a.s.flags + m0, sb + G::FLG, G::T);
        bulk_s2g(a.io.done + m0, sb + G::DONE, G::T);
        bulk_s2g(a.s.steps + e0, sb + G::STEPS, G::E * 4);
        bulk_commit();
      };
      const int pre = n_my < S ? n_my : S;
      for (int j = 0; j < pre; ++j) issue_load(j);
      for (int j = 0; j < n_my; ++j) {
        mbar_wait(bar0 + 8 * (S + j % S), (unsigned)(j / S) & 1u);  // every compute warp has written tile j's outputs
        issue_store(j);
        if (j + S < n_my) {
          bulk_wait_read0();  // the stage has been read out: refill it
          issue_load(j + S);
        }
      }
      bulk_wait_all0();
    }
    return;
  }

  // ================= compute warps: one lane per UAV =================
  Lane L;
  L.N = NT;
  L.lanes_used = G::LANES;
  L.lane = lane;
  L.valid_lanes = G::LANES;
  L.valid = lane < G::LANES;
  const int e_local = L.valid ? lane / NT : 0;
  L.i = L.valid ? lane - e_local * NT : 0;  // idle lanes shadow UAV 0 of the warp's first env and never store
  L.base = e_local * NT;
  L.ring = e_local * (NT + 1);
  L.envmask = NT >= 32 ? 0xffffffffu : ((1u << NT) - 1u);
  const int u_slot = warp * G::LANES + (L.valid ? lane : 0);
  const int e_slot = warp * G::EPW + e_local;
  float* const ring = rings + warp * G::RING_FLOATS;
  const WarpScratch ws{reinterpret_cast<float4*>(ring), reinterpret_cast<float4*>(ring + 192), ring + 320, nullptr};

  for (int j = 0; j < n_my; ++j) {
    const int s = j % S;
    const int tile = (int)blockIdx.x + j * (int)gridDim.x;
    L.warp_m0 = tile * G::T + warp * G::LANES;
    L.m = L.warp_m0 + (L.valid ? lane : 0);
    L.env = tile * G::E + e_slot;
    mbar_wait(bar0 + 8 * s, (unsigned)(j / S) & 1u);  // the tile's inputs have landed in stage s
    SmemIO<G, FINAL> io{a, stages + s * G::STAGE_BYTES, u_slot, e_slot, L.m, L.env, L.valid};
    step_core<NT>(a, ws, L, io);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar0 + 8 * (S + s));
  }
}

}  // namespace uavca
