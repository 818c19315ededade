// uavca_policy.cu — fused acting path of the shared SAC policy ("next" row 2 of SURVEY.md §8f) on the 5th-generation
// tensor cores: obs [M,10] -> relu(W1) -> relu(W2) -> (mean, log_std) -> tanh(mean + exp(log_std) * eps) in ONE kernel.
//
// Restates GaussianPolicy.forward/sample (pytorch_sac_temp/model.py:74-101) as used by SAC.select_action
// (pytorch_sac_temp/sac.py:38-44) for all M = B*N UAVs at once (every UAV shares one policy, test_sac_multi.py:90-91).
// The reference runs it as four cuBLAS GEMVs per UAV per step; batched in eager PyTorch the two hidden layers
// dominate a rollout step (~960 us of 980 us at B=16,384 x N=10 in fp32, 276 us under bf16 autocast, mostly
// elementwise traffic on the [M,256] activations).  Here the activations never leave the SM, all three layers run
// on the tensor cores and the biases ride along as one extra K step against a constant "ones" operand:
//
//   per CTA (persistent, one per SM; 16 warps: warp w owns TMEM lanes 32*(w%4).. and hidden columns 64*(w/4)..),
//   per tile of 128 rows:
//     obs rows -> fp16 A1 [128x16] in smem (10 features, a 1.0 in column 10 that multiplies b1, zero padding)
//     tcgen05.mma  D1[128x256] (TMEM cols 0..255)   = A1 . [W1 | b1]^T                 1 MMA  (M=128, N=256, K=16)
//     tcgen05.ld D1 -> relu -> fp16 -> A2 [128x256] in smem, written straight into the 128B-swizzled K-major
//                                                                         layout the MMA descriptor expects
//     tcgen05.mma  D2[128x256] (TMEM cols 256..511) = A2 . W2^T                          16 MMAs
//     tcgen05.ld D2 -> + b2 -> relu -> the four output heads on the CUDA cores: every thread folds its 64 hidden units
//                   into 4 partial dot products with W3 (packed FP32 FMAs, W3 / b2 broadcast from shared memory), the
//                   four column groups of a row meet in shared memory.  (Round 1 ran the heads as 17 more MMAs of N=16,
//                   which cost as much tensor-pipe time as the 256-wide layer 2, plus a second fp16 round trip of the
//                   activations through shared memory: 2,800 of 6,900 cycles per tile.)
//                  ... layer 1 of the CTA's NEXT tile is issued meanwhile, so that its result is waiting in D1
//     mean, log_std; eps ~ N(0,1) from Philox4x32-10 + Box-Muller (or caller-supplied)
//                   -> action = tanh(mean + exp(clamp(log_std)) eps)
//
// W2 (fp16, 128 KB) stays resident in shared memory for the CTA's lifetime.  The tensor-core operands (W1 | b1, W2,
// activations) are fp16 with fp32 accumulation (11 significand bits, like TF32); b2, W3 and b3 act in fp32: the policy is the learner's side of the
// boundary, not part of the env-step parity contract; tests compare against the fp32 PyTorch policy at 2e-2
// absolute on (mean, log_std) — measured 3e-4.
#include <cuda_fp16.h>

#include <atomic>

#include "uavca_host.h"

namespace uavca {

namespace pol {

constexpr int kRows = 128;      // rows per tile = TMEM lanes
constexpr int kColGroups = 4;   // the 256 hidden columns of a row are split over 4 threads (64 each)
constexpr int kThreads = kRows * kColGroups;  // 16 warps: warp w works on TMEM lanes 32*(w%4).., columns 64*(w/4)..
constexpr int kHidden = 256;
constexpr int kInPad = 16;      // obs features padded to one MMA K step
constexpr int kObs = 10;
constexpr uint32_t kTmemCols = 512;

// shared-memory carve-up (bytes; the swizzled operands need 1024-byte alignment)
constexpr int kW2Bytes = kHidden * kHidden * 2;       // 131072: 4 K-blocks x [256 rows x 128 B], 128B swizzle
constexpr int kA2Bytes = kRows * kHidden * 2;         //  65536: 4 K-blocks x [128 rows x 128 B], 128B swizzle
constexpr int kW1Bytes = kHidden * kInPad * 2;        //   8192: K = 16, no swizzle, 8x16B core matrices ([W1 | b1])
constexpr int kA1Bytes = kRows * kInPad * 2;          //   4096
constexpr int kW3fBytes = kHidden * 16;               //   4096: float4 per hidden unit = its weight in the four heads
constexpr int kB2fBytes = kHidden * 4;                //   1024: linear2.bias in fp32
constexpr int kPartBytes = 3 * kRows * 16;            //   6144: head partials of column groups 1..3 (float4 per row)
constexpr int kOffW2 = 0;
constexpr int kOffA2 = kOffW2 + kW2Bytes;
constexpr int kOffW1 = kOffA2 + kA2Bytes;
constexpr int kOffA1 = kOffW1 + kW1Bytes;
constexpr int kOffW3f = kOffA1 + kA1Bytes;
constexpr int kOffB2f = kOffW3f + kW3fBytes;
constexpr int kOffPart = kOffB2f + kB2fBytes;
constexpr int kOffB3 = kOffPart + kPartBytes;         // float4: the four head biases
constexpr int kOffBar = kOffB3 + 16;                  // mbarriers + tmem base
constexpr int kSmemBytes = kOffBar + 64;
constexpr int kSmemAlloc = kSmemBytes + 1024;         // slack to align the dynamic buffer to 1024 B
static_assert(kSmemAlloc <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors (cute/arch/mma_sm100_desc.hpp layouts, written out) ------------------------------------------------
// shared-memory matrix descriptor: [0,14) start>>4, [16,30) leading byte offset>>4, [32,46) stride byte offset>>4,
// [46,48) version = 1 (sm_100), [61,64) layout type (0 none, 2 = 128B swizzle)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor, kind::f16: c_format F32 (1 << 4), a/b format F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24); }

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate, uint32_t kIdesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "POL_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra POL_DONE_%=;\n"
      "bra POL_WAIT_%=;\n"
      "POL_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float4 tmem_ld4(uint32_t taddr) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// byte offset of the 16-byte chunk holding elements [8c, 8c+8) of row r inside a K-major, 128B-swizzled operand
// whose K-blocks (64 halves = 128 B per row) are `block_bytes` apart: rows 128 B apart in 8-row atoms of 1024 B,
// chunk index XORed with the row inside the atom (Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t sw128_chunk(int r, int c, int block_bytes) {
  return (uint32_t)((c >> 3) * block_bytes + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4));
}
// no-swizzle K-major operand with K = 16 (two 16-byte chunks per row): 8x16B core matrices, the two K chunks of an
// 8-row group 128 B apart (leading byte offset), 8-row groups 256 B apart (stride byte offset)
__device__ __forceinline__ uint32_t k16_chunk(int r, int c) { return (uint32_t)((r >> 3) * 256 + c * 128 + (r & 7) * 16); }

struct Args {
  const float* obs;     // [M][10]
  const __half* w1;     // [256][16]: linear1.weight in columns 0..9, linear1.bias in column 10, zeros
  const __half* w2;     // [256][256]: linear2.weight
  const __half* w2b;    // [256][16]: linear2.bias in column 0, zeros
  const __half* w3;     // [16][256]: rows 0,1 mean_linear.weight, rows 2,3 log_std_linear.weight, zeros
  const __half* w3b;    // [16][16]: the four head biases in column 0 of rows 0..3, zeros
  const float* noise;   // [M][2] standard normal draws, or nullptr -> Philox
  float* action;        // [M][2]
  float* head;          // [M][4] (mean0, mean1, log_std0, log_std1) or nullptr
  long long M;
  unsigned seed_lo, seed_hi;
  unsigned long long ctr;               // Philox counter words 2,3 = ctr + *ctr_dev
  const unsigned long long* ctr_dev;    // nullable
};

__global__ void __launch_bounds__(kThreads, 1) policy_act_kernel(const __grid_constant__ Args a) {
  extern __shared__ unsigned char smem_raw[];
  // aligned to 1024 B by pointer arithmetic on the __shared__ array (keeps the address space: LDS / STS, not generic LD / ST)
  unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q = warp & 3;          // TMEM lane quarter this warp may access
  const int g = warp >> 2;         // column group
  const int r = q * 32 + lane;     // row of the tile (= TMEM lane)
  const uint32_t sbase = smem_u32(sm);
  const uint32_t bar = sbase + kOffBar, bar1 = sbase + kOffBar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 16);

  // ---- one-time setup: TMEM, barrier, weights into shared memory in MMA layouts
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) {
    bar_init(bar, 1);
    bar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    // weights: 16-byte asynchronous copies (LDGSTS) straight into the MMA layouts, all in flight at once
    const uint4* w2g = reinterpret_cast<const uint4*>(a.w2);  // 256 rows x 32 chunks of 8 halves
#pragma unroll 8
    for (int idx = t; idx < kHidden * 32; idx += kThreads) {
      const int n = idx >> 5, c = idx & 31;
      cp_async16(sbase + kOffW2 + sw128_chunk(n, c, kHidden * 128), w2g + idx);
    }
    // [256][16] operand [W1 | b1]: 512 chunks, one per thread
    cp_async16(sbase + kOffW1 + k16_chunk(t >> 1, t & 1), reinterpret_cast<const uint4*>(a.w1) + t);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (t < kHidden) {  // fp32 tables of the CUDA-core head layer: W3 transposed to one float4 per hidden unit, b2
      const __half* w3 = a.w3;
      reinterpret_cast<float4*>(sm + kOffW3f)[t] = make_float4(__half2float(w3[t]), __half2float(w3[kHidden + t]),
                                                               __half2float(w3[2 * kHidden + t]), __half2float(w3[3 * kHidden + t]));
      reinterpret_cast<float*>(sm + kOffB2f)[t] = __half2float(a.w2b[t * kInPad]);
    }
    if (t == 0)
      *reinterpret_cast<float4*>(sm + kOffB3) = make_float4(__half2float(a.w3b[0]), __half2float(a.w3b[kInPad]),
                                                            __half2float(a.w3b[2 * kInPad]), __half2float(a.w3b[3 * kInPad]));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);     // this thread's TMEM lane, column 0
  const uint32_t lane_addr = lane_base + (uint32_t)(g * 64);        // first of its 64 hidden columns
  uint32_t phase = 0;
  const unsigned long long ctr = a.ctr + (a.ctr_dev ? *a.ctr_dev : 0ull);

  const long long tiles = (a.M + kRows - 1) / kRows;
  // column group 0 owns the row's observation (zeros past the end), fetched one tile ahead
  auto load_row = [&](long long tile, float2 (&o)[kObs / 2]) {
    const long long row = tile * kRows + r;
#pragma unroll
    for (int k = 0; k < kObs / 2; ++k) o[k] = make_float2(0.f, 0.f);
    if (g == 0 && tile < tiles && row < a.M) {
      const float2* gp = reinterpret_cast<const float2*>(a.obs + row * kObs);
#pragma unroll
      for (int k = 0; k < kObs / 2; ++k) o[k] = __ldcs(gp + k);
    }
  };
  // accumulator columns [col0, col0+64) of this thread's row -> relu -> fp16 -> K-block g of the A2 operand
  auto relu_to_a2 = [&](uint32_t col0) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(col0 + (uint32_t)(c * 32), v);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {  // four 16-byte chunks of 8 hidden units
        __half2 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(fmaxf(v[k4 * 8 + 2 * j], 0.f), fmaxf(v[k4 * 8 + 2 * j + 1], 0.f));
        *reinterpret_cast<uint4*>(sm + kOffA2 + sw128_chunk(r, g * 8 + c * 4 + k4, kRows * 128)) = *reinterpret_cast<const uint4*>(h);
      }
    }
  };
  // D2 = A2 . W2^T (K = 256 in 16 steps of the 128B-swizzled K-major operands)
  auto issue_layer2 = [&]() {
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // K = 16 halves = 32 bytes per step inside the 128-byte swizzle atom
        const uint64_t da = smem_desc(sbase + kOffA2 + kb * (kRows * 128) + ks * 32, 16, 1024, 2);
        const uint64_t db = smem_desc(sbase + kOffW2 + kb * (kHidden * 128) + ks * 32, 16, 1024, 2);
        mma_f16(tmem + 256u, da, db, (uint32_t)((kb | ks) != 0), idesc(kHidden));
      }
    }
    mma_commit(bar);
  };
  // The output heads of this thread's 64 hidden units: h = relu(D2 + b2), partial dot products with W3 in packed FP32
  // (two heads per FFMA2, W3 / b2 read as warp-wide broadcasts).  Returns (mean0, mean1, log_std0, log_std1) partials.
  auto head_partials = [&]() {
    float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
    const float4* w3f = reinterpret_cast<const float4*>(sm + kOffW3f) + g * 64;
    const float2* b2f = reinterpret_cast<const float2*>(sm + kOffB2f) + g * 32;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(lane_addr + 256u + (uint32_t)(c * 32), v);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float2 hb = __fadd2_rn(make_float2(v[2 * k], v[2 * k + 1]), b2f[c * 16 + k]);
        const float h0 = fmaxf(hb.x, 0.f), h1 = fmaxf(hb.y, 0.f);
        const float4 wa = w3f[c * 32 + 2 * k], wb = w3f[c * 32 + 2 * k + 1];
        acc01 = __ffma2_rn(make_float2(h0, h0), make_float2(wa.x, wa.y), acc01);
        acc23 = __ffma2_rn(make_float2(h0, h0), make_float2(wa.z, wa.w), acc23);
        acc01 = __ffma2_rn(make_float2(h1, h1), make_float2(wb.x, wb.y), acc01);
        acc23 = __ffma2_rn(make_float2(h1, h1), make_float2(wb.z, wb.w), acc23);
      }
    }
    return make_float4(acc01.x, acc01.y, acc23.x, acc23.y);
  };

  // this thread's share of the layer-1 operand of a tile: the observation row as 16 halves (10 features, 1.0 for
  // the bias column, zero padding); column group 0 only
  auto write_a1 = [&](const float2 (&o)[kObs / 2]) {
    if (g == 0) {
      __half2 h[8];
#pragma unroll
      for (int k = 0; k < 5; ++k) h[k] = __floats2half2_rn(o[k].x, o[k].y);
      h[5] = __floats2half2_rn(1.f, 0.f);
      h[6] = h[7] = __floats2half2_rn(0.f, 0.f);
      *reinterpret_cast<uint4*>(sm + kOffA1 + k16_chunk(r, 0)) = *reinterpret_cast<const uint4*>(&h[0]);
      *reinterpret_cast<uint4*>(sm + kOffA1 + k16_chunk(r, 1)) = *reinterpret_cast<const uint4*>(&h[4]);
    }
  };
  auto issue_layer1 = [&]() {
    mma_f16(tmem, smem_desc(sbase + kOffA1, 128, 256, 0), smem_desc(sbase + kOffW1, 128, 256, 0), 0u, idesc(kHidden));
    mma_commit(bar1);
  };

  float2 o_next[kObs / 2];
  load_row(blockIdx.x, o_next);
#ifdef UAVCA_POLICY_TIMING
  long long tk[6] = {0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
  const long long tstart = tprev;
#define POL_TICK(i) { long long now = clock64(); tk[i] += now - tprev; tprev = now; }
#else
#define POL_TICK(i)
#endif
  // prologue: layer 1 of the first tile
  write_a1(o_next);
  load_row((long long)blockIdx.x + gridDim.x, o_next);
  fence_async_smem();
  fence_before();
  __syncthreads();
  if (t == 0 && (long long)blockIdx.x < tiles) {
    fence_after();
    issue_layer1();
  }
  uint32_t phase1 = 0;  // parity of bar1 (layer 1 of a tile), `phase` that of bar (layers 2 and 3)

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long row = tile * kRows + r;
    const bool live = row < a.M;
    const bool has_next = tile + gridDim.x < tiles;

    bar_wait(bar1, phase1);  // D1 = [obs | 1] . [W1 | b1]^T of this tile (issued one step ahead)
    phase1 ^= 1u;
    fence_after();
    POL_TICK(0)

    // ---- layer 1 epilogue -> A2; layer 2 on the tensor cores
    relu_to_a2(lane_addr);
    POL_TICK(1)
    fence_async_smem();
    fence_before();
    __syncthreads();
    POL_TICK(2)
    if (t == 0) {
      fence_after();
      issue_layer2();
    }
    bar_wait(bar, phase);
    phase ^= 1u;
    fence_after();
    POL_TICK(3)

    // ---- layer 2 epilogue = the output heads, on the CUDA cores; the next tile's layer-1 operand
    const float4 part = head_partials();
    if (g != 0) reinterpret_cast<float4*>(sm + kOffPart)[(g - 1) * kRows + r] = part;
    write_a1(o_next);
    load_row(tile + 2 * (long long)gridDim.x, o_next);  // two tiles ahead: in flight for a whole iteration
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (t == 0 && has_next) {  // D1 is free (its epilogue ran at the top of this iteration): layer 1 of the next tile
      fence_after();
      issue_layer1();
    }
    POL_TICK(4)

    // ---- sample and squash (one thread per row: the column-group-0 warps)
    if (g == 0) {
      const float4* pp = reinterpret_cast<const float4*>(sm + kOffPart) + r;
      const float4 p1 = pp[0], p2 = pp[kRows], p3 = pp[2 * kRows], b3 = *reinterpret_cast<const float4*>(sm + kOffB3);
      const float4 hd = make_float4(part.x + p1.x + p2.x + p3.x + b3.x, part.y + p1.y + p2.y + p3.y + b3.y,
                                    part.z + p1.z + p2.z + p3.z + b3.z, part.w + p1.w + p2.w + p3.w + b3.w);
      if (live) {
        const float m0 = hd.x, m1 = hd.y;
        const float l0 = fminf(fmaxf(hd.z, -20.f), 2.f);  // LOG_SIG_MIN / LOG_SIG_MAX (model.py:6-7,79)
        const float l1 = fminf(fmaxf(hd.w, -20.f), 2.f);
        float e0, e1;
        if (a.noise) {
          const float2 z = reinterpret_cast<const float2*>(a.noise)[row];
          e0 = z.x; e1 = z.y;
        } else {  // Box-Muller on one Philox4x32-10 block keyed by (seed, row, call counter)
          const uint4 rn = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32), a.seed_lo, a.seed_hi);
          const float u1 = ((float)(rn.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
          const float u2 = ((float)(rn.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
          const float rad = sqrtf(-2.0f * __logf(u1));
          float sn, cs;
          __sincosf(6.28318530717958647692f * u2, &sn, &cs);
          e0 = rad * cs; e1 = rad * sn;
        }
        const float x0 = fmaf(__expf(l0), e0, m0), x1 = fmaf(__expf(l1), e1, m1);
        reinterpret_cast<float2*>(a.action)[row] = make_float2(tanhf(x0), tanhf(x1));  // model.py:90-91
        if (a.head) reinterpret_cast<float4*>(a.head)[row] = make_float4(m0, m1, l0, l1);
      }
    }
    // the next iteration's first __syncthreads (after its layer-1 epilogue) orders this iteration's TMEM reads of D2 and
    // shared-memory reads of the partials before the next layer-2 MMAs / partial writes
    fence_before();
  }

#ifdef UAVCA_POLICY_TIMING
  POL_TICK(5)
  if (t == 0 && blockIdx.x == 0 && a.head) {
    for (int i = 0; i < 6; ++i) a.head[i] = (float)tk[i];
    a.head[6] = (float)(clock64() - tstart);
    a.head[7] = (float)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  }
#endif
  // ---- teardown
  fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

}  // namespace pol

cudaError_t launch_policy_act(const float* obs, long long M, const void* w1, const void* w2, const void* w2b, const void* w3,
                              const void* w3b, const float* noise, unsigned long long seed, unsigned long long counter,
                              const unsigned long long* counter_dev, float* action, float* head, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  static std::atomic<bool> configured[64];  // per device; setting the attribute twice is harmless, the flag only skips the call
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int sms = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(pol::policy_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pol::kSmemAlloc);
    if (e != cudaSuccess) return e;
    configured[dev].store(true, std::memory_order_release);
  }
  pol::Args a{};
  a.obs = obs;
  a.w1 = reinterpret_cast<const __half*>(w1); a.w2 = reinterpret_cast<const __half*>(w2);
  a.w2b = reinterpret_cast<const __half*>(w2b); a.w3 = reinterpret_cast<const __half*>(w3);
  a.w3b = reinterpret_cast<const __half*>(w3b);
  a.noise = noise; a.action = action; a.head = head; a.M = M;
  a.seed_lo = (unsigned)seed; a.seed_hi = (unsigned)(seed >> 32);
  a.ctr = counter; a.ctr_dev = counter_dev;
  const long long tiles = (M + pol::kRows - 1) / pol::kRows;
  const int grid = (int)(tiles < sms ? tiles : sms);
  pol::policy_act_kernel<<<grid, pol::kThreads, pol::kSmemAlloc, st>>>(a);
  return cudaGetLastError();
}

}  // namespace uavca
