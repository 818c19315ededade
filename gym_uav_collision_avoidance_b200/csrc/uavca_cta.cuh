// uavca_cta.cuh — MultiUAVWorld2D.step for envs of 17..25 UAVs: envs packed across the warps of a CTA.
//
// The warp kernels (uavca_multi.cuh) keep an env inside one warp, so 17 <= N <= 31 runs ONE env per warp and idles
// 32 - N lanes of every instruction (N=20: 37 %, N=24: 25 % — and the reference's own sweep is N = 1..24,
// test_sac_multi_score.py:12,30).  Here a CTA of 128 threads holds floor(128 / N) whole envs back to back (N=24: 5 envs
// on 120 lanes, N=20: 6, N=17: 7), an env may straddle a warp boundary, and everything the warp kernels do with
// shuffles / ballots / __syncwarp goes through shared memory and CTA barriers instead:
//   neighbour sweep   the same pair-packed rings, candidate table and pair_scan / neighbours code (uavca_multi.cuh), per CTA
//   done bits, reach / hard-collision counts of an env   shared-memory atomics, read back after one barrier
//   auto-reset        rare: the warps of the CTA share out the envs that start a new episode, one warp replays
//                     MultiUAVWorld2D.reset for a whole env (reset_multi, as the warp kernels do) and writes the new
//                     state; after a barrier the env's own threads pick their UAV up and observe it
// Per-UAV arithmetic (kinematics, reward, collision / done logic) is the same sequence of intrinsics as step_core.
#pragma once

#include "uavca_multi.cuh"

namespace uavca {

constexpr int kCtaMaxEnvs = 12;  // floor(128 / 11) = 11

// Which env widths the CTA packing serves: where 128 threads hold clearly more envs than 4 warps do — 17..25 UAVs (5..7
// envs against 4) and 11 UAVs (11 envs against 8); 12 UAVs (10 against 8) only pays in the K-steps-per-launch kernel.
// Elsewhere the packing gains at most one env in eight and the CTA barriers cost more than that.  Measured at B*N = 2 Mi
// UAVs: N=11 96.5 -> 89.6 us per step, rollout 80.1 -> 69.6; N=12 90.7 -> 90.7, rollout 73.8 -> 69.5.
inline bool cta_packs(int N, bool rollout) {
  if (N == 11 || (N == 12 && rollout)) return true;
  return N > 16 && kThreads / N >= 5;
}

struct CtaShared {
  float4 pairs[kThreads + kCtaMaxEnvs];  // env e at float4 offset e * (N + 1): sum over envs <= 128 + 7
  float4 cand[kThreads];
  float th[kThreads];
  float stage[kThreads * 10];
  float live[kThreads];                  // rewards[i] * (1 - dones[i]) of every UAV (track_scores)
  unsigned done_bits[kCtaMaxEnvs];
  int reach_inc[kCtaMaxEnvs], coll_inc[kCtaMaxEnvs];
  unsigned reset_env[kCtaMaxEnvs];       // env starts a new episode this step
  unsigned episode[kCtaMaxEnvs];         // its episode counter before the reset
  unsigned bad;
};

// rows [0, rows) of the CTA's staged observations -> g (the CTA's first row in the output tensor)
__device__ __forceinline__ void cta_flush_rows(const float* stage, float* g, int rows, bool aligned16) {
  if (aligned16) {
    const int n4 = rows * 10 / 4;  // rows * 40 bytes is a multiple of 16 when rows is even
    const float4* s4 = reinterpret_cast<const float4*>(stage);
    float4* g4 = reinterpret_cast<float4*>(g);
    for (int k = threadIdx.x; k < n4; k += kThreads) st_stream(g4 + k, s4[k]);
    if ((rows * 10) & 3) {  // odd number of rows: the last 8 bytes
      const int k2 = rows * 5 - 1;
      if (threadIdx.x == 0) st_stream(reinterpret_cast<float2*>(g) + k2, reinterpret_cast<const float2*>(stage)[k2]);
    }
  } else {
    const float2* s2 = reinterpret_cast<const float2*>(stage);
    float2* g2 = reinterpret_cast<float2*>(g);
    for (int k = threadIdx.x; k < rows * 5; k += kThreads) st_stream(g2 + k, s2[k]);
  }
}

// One step of the CTA's envs.  `u` is this thread's UAV (in registers), `act` its action of this step, `row0` the offset of
// this step's rows in the per-UAV outputs (0, or k * B * N inside a rollout), `env0` the same for the per-env output.
// ROLL: the state stays in registers between the K steps of a launch (the caller stores it once, at the end).
template <bool ROLL>
__device__ __forceinline__ void cta_step(const KernelArgs& a, CtaShared& sh, const Lane& L, int e, int envs_here, Uav& u,
                                         float2 act, size_t row0, size_t env0) {
  const Consts& c = a.c;
  const int N = a.N;
  const int t = threadIdx.x;
  const WarpScratch ws{sh.pairs, sh.cand, sh.th, sh.stage};
  const bool leader = L.valid & (L.i == 0);
  const int steps_new = (L.valid ? a.s.steps[L.env] : 0) + 1;
  if (t < kCtaMaxEnvs) { sh.done_bits[t] = 0u; sh.reach_inc[t] = 0; sh.coll_inc[t] = 0; }
  if (t == 0) sh.bad = 0u;
  if (a.io.action_mode != UAVCA_ACTION_CARTESIAN) act = map_action(act, a.io.action_mode, c);

  // ---- per UAV, as step_core: UAVAgent.step (uav_agent.py:23-36), heading / distance, reward shaping (:188-195)
  const bool parked = (u.flags & UAVCA_FLAG_PARKED) != 0u;
  const float ox = u.px, oy = u.py;
  {
    double vx = u.vx, vy = u.vy;
    float px = u.px, py = u.py;
    integrate((double)act.x, (double)act.y, vx, vy, px, py, c);
    if (!parked) { u.vx = vx; u.vy = vy; u.px = px; u.py = py; }
  }
  const Own w = own_features(c, u.px, u.py, u.tx, u.ty, u.vx, u.vy);
  const float dist = parked ? 0.f : w.dist;
  const float prev_d = parked ? 0.f : u.prev;
  float r;
  {
    const float inv_init = rcp_approx(u.init);
    const float m = fminf(c.vm2_f * inv_init, 1.0f);
    r = fmaf(50.0f * c.inv_vm2_f, __fsub_rn(prev_d, dist), -0.01f * m);
    const float q = dist * inv_init * (1.0f / 1.5f);
    r *= (r > 0.0f) ? (1.0f - q) : (1.0f + q);
    r = fmaf(-0.01f * 3.14159274101257324f, fabsf(w.dth_u), r);
  }

  // ---- both pairwise passes (pair_scan / neighbours of uavca_multi.cuh on CTA-wide tables)
  auto publish_cta = [&](float nx, float ny, float oxx, float oyy, float th_u) {
    if (L.valid) {
      float* ring = reinterpret_cast<float*>(sh.pairs + L.ring);
      const int s0 = L.i, s1 = L.i + N;
      const int a0 = ((s0 >> 1) << 2) + (s0 & 1), a1 = ((s1 >> 1) << 2) + (s1 & 1);
      ring[a0] = nx; ring[a0 + 2] = ny;
      ring[a1] = nx; ring[a1 + 2] = ny;
      sh.cand[t] = make_float4(nx, ny, oxx, oyy);
      sh.th[t] = th_u;
    }
    __syncthreads();
  };
  publish_cta(u.px, u.py, ox, oy, w.th_u);
  float smin;
  int k1, k2, t3;
  pair_scan<0>(c, ws, L, u.px, u.py, k1, k2, t3);
  const ObsTail tail = neighbours<0, true>(c, ws, L, u.px, u.py, w.th_u, k1, k2, t3, smin);

  // ---- collisions (:199-210) and done logic (:213-227), as step_core
  const bool collision = smin <= c.s_coll_le;
  r = collision ? -2.0f : r;
  const bool hard = (smin <= c.s_hard_le) & ((u.flags & (UAVCA_FLAG_PARKED | UAVCA_FLAG_COLLIDED)) == 0u) & L.valid;
  if (hard) u.flags |= UAVCA_FLAG_COLLIDED;
  const bool inside = (u.px >= c.lox_f) & (u.px <= c.hix_f) & (u.py >= c.loy_f) & (u.py <= c.hiy_f);
  const bool reached = (parked ? true : (w.ssq < c.s_reach_lt)) & !collision & (w.vsq < c.reach_speed_sq);
  const bool newly_reached = reached & !parked & L.valid;
  const bool done = (reached | (!inside & (a.io.evaluate == 0))) & L.valid;
  double vsq_obs = w.vsq;
  if (reached) {
    u.flags |= UAVCA_FLAG_PARKED;
    const double2 fv = finish_velocity(u.vx, u.vy, w.vsq);
    u.vx = fv.x; u.vy = fv.y;
    r += 10.0f;
    vsq_obs = sq64(u.vx, u.vy);
  }
  u.prev = dist;

  // ---- outputs of this UAV; env-level facts through shared memory
  {
    float2 o01, o23;
    obs_own(c, w, vsq_obs, o01, o23);
    stage_own(sh.stage, t, o01, o23);
    stage_neighbours(sh.stage, t, tail);
  }
  if (L.valid) {
    st_stream(a.io.reward + row0 + L.m, r);
    st_stream(a.io.done + row0 + L.m, (uint8_t)done);
  }
  const bool bad = (!(fabsf(r) <= 3.4e38f) | !(fabsf(u.px) + fabsf(u.py) <= 3.4e38f)) & L.valid;
  if (done) atomicOr(&sh.done_bits[e], 1u << L.i);
  if (newly_reached) atomicAdd(&sh.reach_inc[e], 1);
  if (hard) atomicAdd(&sh.coll_inc[e], 1);
  if (bad) atomicAdd(&sh.bad, 1u);
  if (c.track_scores) sh.live[t] = (done | !L.valid) ? 0.0f : r;
  __syncthreads();

  // ---- per-env bookkeeping: reset decision, counters, scores
  const unsigned done_env = L.valid ? sh.done_bits[e] : 0u;
  const bool rs = (((done_env & c.rs_any_mask) != 0u) | (((done_env ^ L.envmask) | c.rs_all_off) == 0u) |
                   (steps_new >= c.steps_limit)) & L.valid;
  unsigned episode = 0;
  if (leader) {
    if (a.io.reset_mask) a.io.reset_mask[env0 + L.env] = (uint8_t)rs;
    if (c.track_scores) {
      // the same pairwise order as env_sum of the warp kernels (offsets 1, 2, 4, 8, 16), so both give the same float
      float* v = sh.live + L.base;
      for (int off = 1; off < 32; off <<= 1)
        for (int i = 0; i < N; ++i) v[i] += (i + off < N) ? v[i + off] : 0.0f;  // ascending i: v[i + off] is still last round's
      const float live = v[0];
      double2 sc = a.s.score[L.env];
      sc.x += (double)r;
      sc.y += (double)live;
      a.s.score[L.env] = sc;
    }
    const int reach_inc = sh.reach_inc[e], coll_inc = sh.coll_inc[e];
    if (rs) {
      episode = a.s.episode[L.env];
      if (episode > 0u) {  // fold the finished episode into the running totals
        atomicAdd(a.s.stats + 0, 1ull);
        atomicAdd(a.s.stats + 1, (unsigned long long)(a.s.reach[L.env] + reach_inc));
        atomicAdd(a.s.stats + 2, (unsigned long long)(a.s.coll[L.env] + coll_inc));
        atomicAdd(a.s.stats + 3, (unsigned long long)steps_new);
        if (c.track_scores) fold_scores(a.s, L.env);
      }
      a.s.steps[L.env] = 0; a.s.reach[L.env] = 0; a.s.coll[L.env] = 0;  // :166-168
      a.s.episode[L.env] = episode + 1u;
      sh.episode[e] = episode;
    } else {
      a.s.steps[L.env] = steps_new;
      if (reach_inc) a.s.reach[L.env] += reach_inc;  // :221
      if (coll_inc) a.s.coll[L.env] += coll_inc;     // :209
    }
  }
  if (leader) sh.reset_env[e] = rs;
  if (t == 0 && sh.bad) atomicAdd(a.s.stats + 6, (unsigned long long)sh.bad);
  float* const g_obs = a.io.obs + (row0 + (size_t)L.warp_m0) * UAVCA_OBS_DIM_MULTI;
  const bool aligned16 = ((L.lanes_used * (int)blockIdx.x) & 1) == 0;  // this CTA's first row starts on a 16-byte boundary
  if (a.io.final_obs) cta_flush_rows(sh.stage, a.io.final_obs + (row0 + (size_t)L.warp_m0) * UAVCA_OBS_DIM_MULTI, L.valid_lanes, aligned16);

  if (!__syncthreads_or(rs)) {  // nobody resets: the common case
    cta_flush_rows(sh.stage, g_obs, L.valid_lanes, aligned16);
    if (!ROLL) store_uav(a.s, L, u, false);
    return;
  }

  // ---- rare: some env of this CTA starts a new episode in place.  Warp w replays the reference's reset for envs
  // w, w + 4 of the CTA (lane i = UAV i) and writes the new state; the barrier makes it visible to the env's own threads.
  if (!ROLL && !rs) store_uav(a.s, L, u, false);
  for (int e2 = t >> 5; e2 < envs_here; e2 += kWarpsPerBlock) {
    if (!sh.reset_env[e2]) continue;  // warp-uniform
    const int ln = t & 31;
    Lane R;
    R.N = N; R.lanes_used = N; R.lane = ln; R.valid = ln < N; R.i = R.valid ? ln : 0; R.base = 0; R.ring = 0;
    R.envmask = L.envmask; R.env = blockIdx.x * (kThreads / N) + e2; R.warp_m0 = R.env * N; R.valid_lanes = N; R.m = R.warp_m0 + ln;
    Uav nu = u;
    reset_multi(a, R, R.valid, sh.episode[e2], nu);
    store_uav(a.s, R, nu, true);
  }
  __syncthreads();
  if (rs) {
    const float2 p = __ldcg(a.s.pos + L.m), tg = __ldcg(a.s.tgt + L.m);
    const double2 v = __ldcg(a.s.vel + L.m);
    u.px = p.x; u.py = p.y; u.tx = tg.x; u.ty = tg.y; u.vx = v.x; u.vy = v.y;
    u.init = __ldcg(a.s.init + L.m); u.prev = __ldcg(a.s.prev + L.m); u.flags = __ldcg(a.s.flags + L.m);
  }
  const Own w2 = own_features(c, u.px, u.py, u.tx, u.ty, u.vx, u.vy);
  publish_cta(u.px, u.py, u.px, u.py, w2.th_u);
  float unused;
  pair_scan<0>(c, ws, L, u.px, u.py, k1, k2, t3);
  const ObsTail tail2 = neighbours<0, true>(c, ws, L, u.px, u.py, w2.th_u, k1, k2, t3, unused);
  if (rs) {  // rows of the other envs keep this step's observation
    float2 o01, o23;
    obs_own(c, w2, w2.vsq, o01, o23);
    stage_own(sh.stage, t, o01, o23);
    stage_neighbours(sh.stage, t, tail2);
  }
  __syncthreads();
  cta_flush_rows(sh.stage, g_obs, L.valid_lanes, aligned16);
}

// this thread's place in the CTA's packing: env e = t / N of the CTA, UAV t - e N
__device__ __forceinline__ Lane cta_lane(const KernelArgs& a, int& e, int& envs_here) {
  const int N = a.N;
  const int E = kThreads / N;  // envs per CTA
  const int t = threadIdx.x;
  e = t / N;
  Lane L;
  L.N = N;
  L.lanes_used = E * N;
  L.lane = t;                  // index into the CTA's cand / th / stage arrays
  L.i = t - e * N;
  L.base = e * N;
  L.ring = e * (N + 1);
  L.envmask = (1u << N) - 1u;  // N <= 31
  L.env = blockIdx.x * E + e;
  L.warp_m0 = blockIdx.x * E * N;  // first UAV of the CTA
  envs_here = min(E, a.B - blockIdx.x * E);
  L.valid_lanes = envs_here * N;
  L.valid = t < L.valid_lanes;
  L.m = L.warp_m0 + t;
  if (!L.valid) { L.base = 0; L.ring = 0; L.i = 0; }  // idle threads shadow UAV 0 of the CTA's first env and never store
  return L;
}

__global__ void __launch_bounds__(kThreads, 6) step_multi_cta_kernel(const __grid_constant__ KernelArgs a) {
  __shared__ __align__(16) CtaShared sh;
  int e, envs_here;
  const Lane L = cta_lane(a, e, envs_here);
  cudaGridDependencySynchronize();
  Uav u = load_uav(a.s, L);
  const float2 act = L.valid ? ld_stream(a.io.action + L.m) : make_float2(0.f, 0.f);
  cudaTriggerProgrammaticLaunchCompletion();
  cta_step<false>(a, sh, L, e, envs_here, u, act, 0, 0);
}

// K steps per launch for the same packing (uavca_rollout): the UAV stays in registers, the per-step outputs go to slot k
// of the [K][...] blocks; actions from the block or the Philox stream, as rollout_multi_kernel draws them.
__global__ void __launch_bounds__(kThreads, 6) rollout_multi_cta_kernel(const __grid_constant__ KernelArgs a,
                                                                         const __grid_constant__ RolloutArgs r) {
  __shared__ __align__(16) CtaShared sh;
  int e, envs_here;
  const Lane L = cta_lane(a, e, envs_here);
  cudaGridDependencySynchronize();
  Uav u = load_uav(a.s, L);
  cudaTriggerProgrammaticLaunchCompletion();
  const long long env_global = a.c.env_base + L.env;
  uint4 words = make_uint4(0u, 0u, 0u, 0u);
  for (int k = 0; k < r.K; ++k) {
    float2 act = make_float2(0.f, 0.f);
    if (r.action_block != nullptr) {
      if (L.valid) act = ld_stream(r.action_block + (size_t)k * r.M + L.m);
    } else {
      const unsigned long long t = r.step0 + (unsigned long long)k;
      if (k == 0 || (t & 1ull) == 0ull) words = action_words(r.seed_lo, r.seed_hi, env_global, L.i, t);
      act = action_from_words(words, t);
      if (r.action_out != nullptr && L.valid) st_stream(r.action_out + (size_t)k * r.M + L.m, act);
    }
    cta_step<true>(a, sh, L, e, envs_here, u, act, (size_t)k * r.M, (size_t)k * r.B);
    __syncthreads();  // the tables and the staged rows of this step are free again
  }
  store_uav(a.s, L, u, true);
}

}  // namespace uavca
