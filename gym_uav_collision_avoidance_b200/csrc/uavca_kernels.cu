// uavca_kernels.cu — sm_100a kernels of the batched UAV collision-avoidance env step and their launchers.
//
//   step_multi_kernel   MultiUAVWorld2D.step      (multi_uav_world_2d.py:177-241, uav_agent.py:23-64), fused with
//                       observation build (:60-109), counters, done logic and in-place auto-reset (:116-175)
//   step_single_kernel  UAVWorld2D.step/reset     (uav_world_2d.py:137-173, 119-135)
//   reset_*/observe_*   standalone reset() and _get_obs()
//
// One pass over HBM per step: every state field and every output is read/written exactly once with coalesced,
// streaming accesses; neighbour interaction stays in registers (warp shuffles).  No tensor cores: there is no
// dense contraction anywhere in the step.
#include "uavca_host.h"
#include "uavca_multi.cuh"

namespace uavca {

// ================================================================================================================
// Multi-UAV world
// ================================================================================================================

template <int NT>
__global__ void __launch_bounds__(kThreads, kMinBlocksPerSM) step_multi_kernel(const __grid_constant__ KernelArgs a) {
  __shared__ __align__(16) float smem[kWarpsPerBlock * kScratchFloats];
  const WarpScratch ws = warp_scratch(smem);
  const Consts& c = a.c;
  const Lane L = make_lane<NT>(a.B, a.N);
  const bool leader = L.valid & (L.i == 0);

  // Programmatic dependent launch: this grid may have been scheduled while the previous kernel of the stream was
  // still draining; everything above overlapped with it, nothing below may (it reads memory that kernel wrote).
  cudaGridDependencySynchronize();

  // every global load of the step is issued here, back to back
  Uav u = load_uav(a.s, L);
  float2 act = make_float2(0.f, 0.f);
  if (L.valid) act = ld_stream(a.io.action + L.m);
  int steps_new = 0;
  if (leader) steps_new = ld_stream(a.s.steps + L.env) + 1;  // :238
  cudaTriggerProgrammaticLaunchCompletion();
  act = map_action(act, a.io.action_mode, c);

  const bool parked = (u.flags & UAVCA_FLAG_PARKED) != 0u;
  const float ox = u.px, oy = u.py;  // position before this step

  // ---- UAVAgent.step (uav_agent.py:23-36); parked UAVs do not move and report (0, 0)
  float dist, prev_d;
  Own w;
  if (__any_sync(kFull, parked)) {  // warp-uniform: most warps hold no parked UAV and skip the selects
    double vx = u.vx, vy = u.vy;
    float px = u.px, py = u.py;
    integrate((double)act.x, (double)act.y, vx, vy, px, py, c);
    if (!parked) { u.vx = vx; u.vy = vy; u.px = px; u.py = py; }
    w = own_features(c, u);  // heading, heading error to the target, distance, |v|^2 (multi_uav_world_2d.py:184-186)
    dist = parked ? 0.f : w.dist;
    prev_d = parked ? 0.f : u.prev;
  } else {
    integrate((double)act.x, (double)act.y, u.vx, u.vy, u.px, u.py, c);
    w = own_features(c, u);
    dist = w.dist;
    prev_d = u.prev;
  }
  publish(ws, L, u.px, u.py, ox, oy, w.th_u);

  // ---- reward shaping (:188-195).  Output only: float32 arithmetic, well inside the 1e-5 tolerance.
  float r;
  {
    const float m = (u.init <= c.vm2_floor_f) ? 1.0f : __fdividef(c.vm2_f, u.init);  // min(vm2/init, 1)
    r = fmaf(50.0f * c.inv_vm2_f, __fsub_rn(prev_d, dist), -0.01f * m);
    const float q = __fdividef(dist, 1.5f * u.init);
    r *= (r > 0.0f) ? (1.0f - q) : (1.0f + q);
    r = fmaf(-0.01f * 3.14159274101257324f, fabsf(w.dth_u), r);
  }

  // ---- both pairwise passes in one sweep over the env's UAVs
  float smin;
  Top2 t;
  pair_scan<NT>(ws, L, u.px, u.py, smin, t);

  // ---- collisions (:199-210), decided in squared-distance space
  const bool in_range = smin < c.s_dsense_lt;
  const bool collision = in_range & (smin <= c.s_two_r_le);
  r = collision ? -2.0f : r;
  const bool hard = in_range & (smin <= c.s_two_h_le) & !parked & ((u.flags & UAVCA_FLAG_COLLIDED) == 0u);
  if (hard) u.flags |= UAVCA_FLAG_COLLIDED;

  // ---- done logic (:213-227)
  const bool slow = w.vsq < c.reach_speed_sq;
  const bool inside = (u.px >= c.lox_f) & (u.px <= c.hix_f) & (u.py >= c.loy_f) & (u.py <= c.hiy_f);
  const bool reached = (dist < c.reach_dist) & !collision & slow;
  const bool newly_reached = reached & !parked;
  bool done = reached | (!inside & (a.io.evaluate == 0));
  if (reached) {  // UAVAgent.finish (uav_agent.py:38-42)
    u.flags |= UAVCA_FLAG_PARKED;
    const double nv = sqrt(w.vsq);
    double fx = __dmul_rn(__ddiv_rn(u.vx, nv), 0.001), fy = __dmul_rn(__ddiv_rn(u.vy, nv), 0.001);
    if ((fx != fx) | (fy != fy)) { fx = 0.0; fy = 0.0; }
    u.vx = fx; u.vy = fy;
    r += 10.0f;
  }
  const double vsq_obs = reached ? sq64(u.vx, u.vy) : w.vsq;
  u.prev = dist;  // :229
  if (!L.valid) done = false;

  // ---- per-env bookkeeping: counters, reset decision
  const unsigned done_env = (__ballot_sync(kFull, done) >> L.base) & L.envmask;
  const int reach_inc = __popc((__ballot_sync(kFull, newly_reached & L.valid) >> L.base) & L.envmask);
  const int coll_inc = __popc((__ballot_sync(kFull, hard & L.valid) >> L.base) & L.envmask);
  steps_new = __shfl_sync(kFull, steps_new, L.base);
  bool rs = false;
  if (c.reset_mode & UAVCA_RESET_ON_DONE0) rs |= (done_env & 1u) != 0u;
  if (c.reset_mode & UAVCA_RESET_ON_ALL_DONE) rs |= done_env == L.envmask;
  if (c.reset_mode & UAVCA_RESET_ON_ANY_DONE) rs |= done_env != 0u;
  if (c.max_steps > 0) rs |= steps_new >= c.max_steps;
  rs &= L.valid;

  if (L.valid) {
    st_stream(a.io.reward + L.m, r);
    st_stream(a.io.done + L.m, (uint8_t)done);
  }
  if (leader && a.io.reset_mask) a.io.reset_mask[L.env] = (uint8_t)rs;

  // ---- observation (:233-235): every UAV at its new position
  float o[10];
  obs_multi(c, ws, L, u.px, u.py, w.th_u, w.dth_u, w.dist, vsq_obs, t, o);

  if (__any_sync(kFull, rs)) {
    // at least one env of this warp starts a new episode in place
    if (a.io.final_obs) store_obs_rows(ws.stage, a.io.final_obs, L, o);
    unsigned episode = 0;
    if (leader) episode = a.s.episode[L.env];
    episode = __shfl_sync(kFull, episode, L.base);
    if (leader) {
      if (rs) {
        if (episode > 0u) {  // fold the finished episode into the running totals
          atomicAdd(a.s.stats + 0, 1ull);
          atomicAdd(a.s.stats + 1, (unsigned long long)(a.s.reach[L.env] + reach_inc));
          atomicAdd(a.s.stats + 2, (unsigned long long)(a.s.coll[L.env] + coll_inc));
          atomicAdd(a.s.stats + 3, (unsigned long long)steps_new);
        }
        a.s.steps[L.env] = 0; a.s.reach[L.env] = 0; a.s.coll[L.env] = 0;  // :166-168
        a.s.episode[L.env] = episode + 1u;
      } else {
        a.s.steps[L.env] = steps_new;
        if (reach_inc) a.s.reach[L.env] += reach_inc;
        if (coll_inc) a.s.coll[L.env] += coll_inc;
      }
    }
    Uav nu = u;
    reset_multi(a, L, rs, episode, nu);
    float no[10];
    observe_state<NT>(c, ws, L, nu, no);
    if (rs) {
      u = nu;
#pragma unroll
      for (int k = 0; k < 10; ++k) o[k] = no[k];
    }
    store_obs_rows(ws.stage, a.io.obs, L, o);
    store_uav(a.s, L, u, rs);
  } else {
    if (leader) {
      a.s.steps[L.env] = steps_new;
      if (reach_inc) a.s.reach[L.env] += reach_inc;  // :221
      if (coll_inc) a.s.coll[L.env] += coll_inc;     // :209
    }
    store_obs_rows(ws.stage, a.io.obs, L, o);
    if (a.io.final_obs) store_obs_rows(ws.stage, a.io.final_obs, L, o);
    store_uav(a.s, L, u, false);
  }
}

template <int NT>
__global__ void __launch_bounds__(kThreads) reset_multi_kernel(const __grid_constant__ KernelArgs a, const uint8_t* mask) {
  __shared__ __align__(16) float smem[kWarpsPerBlock * kScratchFloats];
  const WarpScratch ws = warp_scratch(smem);
  const Lane L = make_lane<NT>(a.B, a.N);
  Uav u = load_uav(a.s, L);
  const bool rs = L.valid && (mask == nullptr || mask[L.env] != 0);
  const bool leader = L.valid & (L.i == 0);
  unsigned episode = 0;
  if (leader) episode = a.s.episode[L.env];
  episode = __shfl_sync(kFull, episode, L.base);
  if (leader && rs) {
    if (episode > 0u) {
      atomicAdd(a.s.stats + 0, 1ull);
      atomicAdd(a.s.stats + 1, (unsigned long long)a.s.reach[L.env]);
      atomicAdd(a.s.stats + 2, (unsigned long long)a.s.coll[L.env]);
      atomicAdd(a.s.stats + 3, (unsigned long long)a.s.steps[L.env]);
    }
    a.s.steps[L.env] = 0; a.s.reach[L.env] = 0; a.s.coll[L.env] = 0;
    a.s.episode[L.env] = episode + 1u;
  }
  reset_multi(a, L, rs, episode, u);
  float o[10];
  observe_state<NT>(a.c, ws, L, u, o);
  if (a.io.obs) {
    if (mask == nullptr) {
      store_obs_rows(ws.stage, a.io.obs, L, o);
    } else if (rs) {  // rows of other envs stay untouched
#pragma unroll
      for (int k = 0; k < 10; ++k) a.io.obs[(size_t)L.m * 10 + k] = o[k];
    }
  }
  if (rs) store_uav(a.s, L, u, true);
}

template <int NT>
__global__ void __launch_bounds__(kThreads) observe_multi_kernel(const __grid_constant__ KernelArgs a) {
  __shared__ __align__(16) float smem[kWarpsPerBlock * kScratchFloats];
  const WarpScratch ws = warp_scratch(smem);
  const Lane L = make_lane<NT>(a.B, a.N);
  const Uav u = load_uav(a.s, L);
  float o[10];
  observe_state<NT>(a.c, ws, L, u, o);
  store_obs_rows(ws.stage, a.io.obs, L, o);
}

// ================================================================================================================
// Single-UAV world: one thread per env
// ================================================================================================================

struct SingleEnv {
  float px, py, tx, ty, init, prev;
  double vx, vy;
  int steps;
};

__device__ __forceinline__ void obs_single(const Consts& c, const SingleEnv& e, bool vel_is_f32, float o[4]) {
  // UAVWorld2D._get_obs (uav_world_2d.py:88-97)
  const float speed = vel_is_f32 ? n32((float)e.vx, (float)e.vy) : sqrt_approx((float)sq64(e.vx, e.vy));
  o[0] = speed * c.inv_vmax_f;
  o[1] = fast_atan2((float)e.vy, (float)e.vx) * c.inv_pi;
  const float tdx = __fsub_rn(e.tx, e.px), tdy = __fsub_rn(e.ty, e.py);
  o[2] = n32(tdx, tdy) * c.inv_diag;
  o[3] = rel_angle((double)tdx, (double)tdy, e.vx, e.vy) * c.inv_pi;
}

__device__ __forceinline__ void reset_single(const KernelArgs& a, long long b, unsigned episode, SingleEnv& e) {
  // UAVWorld2D.reset (uav_world_2d.py:121-131)
  const Consts& c = a.c;
  const long long env_global = c.env_base + b;
  if (c.reset_source == UAVCA_SOURCE_POOL && a.pool.pos != nullptr) {
    const long long p = (env_global + (long long)episode) % a.pool_envs;
    float2 pp = a.pool.pos[p], pt = a.pool.tgt[p];
    double2 pv = a.pool.vel[p];
    e.px = pp.x; e.py = pp.y; e.tx = pt.x; e.ty = pt.y; e.vx = pv.x; e.vy = pv.y;
    e.init = a.pool.init[p]; e.prev = a.pool.prev[p];
  } else {
    float2 p = draw_pair(c, env_global, episode, kStreamPos, 0, 0u, c.lox, c.hix, c.loy, c.hiy);
    float2 v = draw_pair(c, env_global, episode, kStreamVel, 0, 0u, -c.vmax, c.vmax, -c.vmax, c.vmax);
    float2 t = draw_pair(c, env_global, episode, kStreamTgt, 0, 0u, c.lox, c.hix, c.loy, c.hiy);
    e.px = p.x; e.py = p.y; e.vx = (double)v.x; e.vy = (double)v.y; e.tx = t.x; e.ty = t.y;
    e.init = n32(__fsub_rn(e.tx, e.px), __fsub_rn(e.ty, e.py));
    e.prev = e.init;
  }
  e.steps = 0;
}

__device__ __forceinline__ SingleEnv load_single(const StateView& s, long long b) {
  SingleEnv e;
  float2 p = ld_stream(s.pos + b), t = ld_stream(s.tgt + b);
  double2 v = ld_stream(s.vel + b);
  e.px = p.x; e.py = p.y; e.tx = t.x; e.ty = t.y; e.vx = v.x; e.vy = v.y;
  e.init = ld_stream(s.init + b); e.prev = ld_stream(s.prev + b); e.steps = ld_stream(s.steps + b);
  return e;
}

__device__ __forceinline__ void store_single(const StateView& s, long long b, const SingleEnv& e, bool with_target) {
  st_stream(s.pos + b, make_float2(e.px, e.py));
  st_stream(s.vel + b, make_double2(e.vx, e.vy));
  st_stream(s.prev + b, e.prev);
  st_stream(s.steps + b, e.steps);
  if (with_target) {
    st_stream(s.tgt + b, make_float2(e.tx, e.ty));
    st_stream(s.init + b, e.init);
  }
}

__device__ __forceinline__ void fold_single(const KernelArgs& a, long long b, unsigned episode, int steps) {
  if (episode > 0u) {
    atomicAdd(a.s.stats + 0, 1ull);
    atomicAdd(a.s.stats + 1, (unsigned long long)a.s.reach[b]);
    atomicAdd(a.s.stats + 3, (unsigned long long)steps);
  }
  a.s.reach[b] = 0; a.s.coll[b] = 0;
  a.s.episode[b] = episode + 1u;
}

__global__ void __launch_bounds__(kThreads) step_single_kernel(const __grid_constant__ KernelArgs a) {
  const Consts& c = a.c;
  const long long b = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (b >= a.B) return;
  SingleEnv e = load_single(a.s, b);
  const float2 act = map_action(ld_stream(a.io.action + b), a.io.action_mode, c);

  // UAVWorld2D.step (uav_world_2d.py:142-147)
  if (c.single_f32_first_step && e.steps == 0) {
    // float32 action against the float32 reset speed: the quotient is formed in float32 (:122,:142)
    const double qx = (double)__fdiv_rn(__fsub_rn(act.x, (float)e.vx), c.tau_f);
    const double qy = (double)__fdiv_rn(__fsub_rn(act.y, (float)e.vy), c.tau_f);
    const double dvx = clipd(qx, -c.amax, c.amax), dvy = clipd(qy, -c.amax, c.amax);
    e.vx = clipd(__dadd_rn(e.vx, __dmul_rn(dvx, c.tau)), -c.vmax, c.vmax);
    e.vy = clipd(__dadd_rn(e.vy, __dmul_rn(dvy, c.tau)), -c.vmax, c.vmax);
    e.px = __double2float_rn(__dadd_rn((double)e.px, __dmul_rn(e.vx, c.tau)));
    e.py = __double2float_rn(__dadd_rn((double)e.py, __dmul_rn(e.vy, c.tau)));
  } else {
    integrate((double)act.x, (double)act.y, e.vx, e.vy, e.px, e.py, c);
  }
  const float tdx = __fsub_rn(e.tx, e.px), tdy = __fsub_rn(e.ty, e.py);
  const float dist = n32(tdx, tdy);                                               // :150
  const float dth = rel_angle((double)tdx, (double)tdy, e.vx, e.vy);              // :155-156
  // reward (:152-157): float32 under NumPy 2, the angle term is a python float rounded to float32 first
  float r = __fsub_rn(0.0f, __fdiv_rn(1.0f, e.init));
  r = __fadd_rn(r, __fmul_rn(10.0f, __fsub_rn(e.prev, dist)));
  r = __fsub_rn(r, (float)(0.1 * (double)fabsf(dth)));
  const bool inside = (e.px >= c.lox_f) & (e.px <= c.hix_f) & (e.py >= c.loy_f) & (e.py <= c.hiy_f);
  const bool reached = dist < c.reach_dist;                                       // :159
  if (reached) r = __fadd_rn(r, 1000.0f);                                         // :161
  const bool done = reached | !inside;                                            // :159-166
  e.steps += 1;                                                                   // :170
  e.prev = dist;                                                                  // :172
  if (reached) a.s.reach[b] += 1;

  float o[4];
  obs_single(c, e, false, o);
  st_stream(a.io.reward + b, r);
  st_stream(a.io.done + b, (uint8_t)done);
  if (a.io.distance) st_stream(a.io.distance + b, dist);                          // info["distance"] :169
  if (a.io.final_obs) st_stream(reinterpret_cast<float4*>(a.io.final_obs) + b, make_float4(o[0], o[1], o[2], o[3]));

  bool rs = false;
  if (c.reset_mode & (UAVCA_RESET_ON_DONE0 | UAVCA_RESET_ON_ALL_DONE | UAVCA_RESET_ON_ANY_DONE)) rs |= done;
  if (c.max_steps > 0) rs |= e.steps >= c.max_steps;
  if (a.io.reset_mask) a.io.reset_mask[b] = (uint8_t)rs;
  if (rs) {
    const unsigned episode = a.s.episode[b];
    fold_single(a, b, episode, e.steps);
    reset_single(a, b, episode, e);
    obs_single(c, e, true, o);
  }
  st_stream(reinterpret_cast<float4*>(a.io.obs) + b, make_float4(o[0], o[1], o[2], o[3]));
  store_single(a.s, b, e, rs);
}

__global__ void __launch_bounds__(kThreads) reset_single_kernel(const __grid_constant__ KernelArgs a, const uint8_t* mask) {
  const long long b = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (b >= a.B) return;
  if (mask != nullptr && mask[b] == 0) return;
  SingleEnv e = load_single(a.s, b);
  const unsigned episode = a.s.episode[b];
  fold_single(a, b, episode, e.steps);
  reset_single(a, b, episode, e);
  a.s.flags[b] = 0;
  float o[4];
  obs_single(a.c, e, true, o);
  if (a.io.obs) st_stream(reinterpret_cast<float4*>(a.io.obs) + b, make_float4(o[0], o[1], o[2], o[3]));
  store_single(a.s, b, e, true);
}

__global__ void __launch_bounds__(kThreads) observe_single_kernel(const __grid_constant__ KernelArgs a) {
  const long long b = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (b >= a.B) return;
  const SingleEnv e = load_single(a.s, b);
  float o[4];
  obs_single(a.c, e, e.steps == 0, o);
  st_stream(reinterpret_cast<float4*>(a.io.obs) + b, make_float4(o[0], o[1], o[2], o[3]));
}

// ================================================================================================================
// Small utility kernels
// ================================================================================================================

__global__ void __launch_bounds__(kThreads) map_action_kernel(const __grid_constant__ Consts c, const float2* in,
                                                              float2* out, long long M, int mode) {
  const long long m = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (m < M) out[m] = map_action(in[m], mode, c);
}

// out8: finished-episode totals (episodes, reach, collisions, steps), in-flight sums (reach, collisions, steps), B
__global__ void __launch_bounds__(kThreads) stats_kernel(StateView s, int B, long long* out8) {
  long long reach = 0, coll = 0, steps = 0;
  for (long long b = (long long)blockIdx.x * kThreads + threadIdx.x; b < B; b += (long long)gridDim.x * kThreads) {
    reach += s.reach[b]; coll += s.coll[b]; steps += s.steps[b];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    reach += __shfl_xor_sync(kFull, reach, d);
    coll += __shfl_xor_sync(kFull, coll, d);
    steps += __shfl_xor_sync(kFull, steps, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(out8 + 4), (unsigned long long)reach);
    atomicAdd(reinterpret_cast<unsigned long long*>(out8 + 5), (unsigned long long)coll);
    atomicAdd(reinterpret_cast<unsigned long long*>(out8 + 6), (unsigned long long)steps);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out8[0] = (long long)s.stats[0]; out8[1] = (long long)s.stats[1];
    out8[2] = (long long)s.stats[2]; out8[3] = (long long)s.stats[3];
    out8[7] = B;
  }
}

// ================================================================================================================
// Launchers
// ================================================================================================================

static inline int multi_grid(int B, int N) {
  const int epw = 32 / N;
  const long long warps = ((long long)B + epw - 1) / epw;
  return (int)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
}

#define UAVCA_DISPATCH_N(N, CALL)            \
  switch (N) {                               \
    case 2: { constexpr int NT = 2; CALL; } break;   \
    case 4: { constexpr int NT = 4; CALL; } break;   \
    case 5: { constexpr int NT = 5; CALL; } break;   \
    case 8: { constexpr int NT = 8; CALL; } break;   \
    case 10: { constexpr int NT = 10; CALL; } break; \
    case 16: { constexpr int NT = 16; CALL; } break; \
    case 32: { constexpr int NT = 32; CALL; } break; \
    default: { constexpr int NT = 0; CALL; } break;  \
  }

// The step kernel is launched with programmatic stream serialization (PDL): its blocks may become resident while
// the previous kernel in the stream drains, and wait in cudaGridDependencySynchronize() before touching memory.
template <int NT>
static cudaError_t launch_step_multi_n(const KernelArgs& a, int grid, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, step_multi_kernel<NT>, a);
}

cudaError_t launch_step_multi(const KernelArgs& a, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  const int grid = multi_grid(a.B, a.N);
  cudaError_t e = cudaSuccess;
  UAVCA_DISPATCH_N(a.N, (e = launch_step_multi_n<NT>(a, grid, st)));
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_reset_multi(const KernelArgs& a, const uint8_t* mask, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  const int grid = multi_grid(a.B, a.N);
  UAVCA_DISPATCH_N(a.N, (reset_multi_kernel<NT><<<grid, kThreads, 0, st>>>(a, mask)));
  return cudaGetLastError();
}

cudaError_t launch_observe_multi(const KernelArgs& a, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  const int grid = multi_grid(a.B, a.N);
  UAVCA_DISPATCH_N(a.N, (observe_multi_kernel<NT><<<grid, kThreads, 0, st>>>(a)));
  return cudaGetLastError();
}

static inline int flat_grid(long long n) { return (int)((n + kThreads - 1) / kThreads); }

cudaError_t launch_step_single(const KernelArgs& a, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  step_single_kernel<<<flat_grid(a.B), kThreads, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_reset_single(const KernelArgs& a, const uint8_t* mask, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  reset_single_kernel<<<flat_grid(a.B), kThreads, 0, st>>>(a, mask);
  return cudaGetLastError();
}

cudaError_t launch_observe_single(const KernelArgs& a, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  observe_single_kernel<<<flat_grid(a.B), kThreads, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_map_action(const Consts& c, const float* in, float* out, long long M, int mode, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  map_action_kernel<<<flat_grid(M), kThreads, 0, st>>>(c, reinterpret_cast<const float2*>(in),
                                                       reinterpret_cast<float2*>(out), M, mode);
  return cudaGetLastError();
}

cudaError_t launch_stats(const StateView& s, int B, long long* out8, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out8, 0, 8 * sizeof(long long), st);
  if (e != cudaSuccess) return e;
  int grid = flat_grid(B);
  if (grid > 148 * 4) grid = 148 * 4;
  if (grid < 1) grid = 1;
  stats_kernel<<<grid, kThreads, 0, st>>>(s, B, out8);
  return cudaGetLastError();
}

}  // namespace uavca
