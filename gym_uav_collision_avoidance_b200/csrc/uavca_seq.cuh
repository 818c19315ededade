// uavca_seq.cuh — the general multi-UAV step: one thread per env, UAVs in sequence, templated on the position type.
//
// Two corners of MultiUAVWorld2D fall outside the warp-per-env kernels of uavca_multi.cuh and run here instead:
//   * P = double: episodes started by reset(circular=True).  The reference assigns float64 arrays to `location` and
//     `target_location` there (multi_uav_world_2d.py:157-163), so `self.location += dx` (uav_agent.py:28-29), every
//     np.linalg.norm and every threshold test of such an episode is float64.  This path keeps the positions, targets,
//     init / prev distances in float64 (state fields pos64 / tgt64 / init64 / prev64) and follows that dtype sequence:
//     flags, positions and velocities come out bit-identical to the reference (tests/golden/circular_*.npz).
//   * P = float with N > 32: more UAVs per env than a warp holds (`num_agents` is unbounded in the reference,
//     multi_uav_world_2d.py:13,36-41; its scripts stop at 24).  Same float32 semantics as the warp kernels.
// Neither is a throughput path (circular episodes are the plotting / evaluation layout of
// test_sac_multi_plot_trajectory.py; N > 32 appears in no reference script): the code mirrors the reference's
// sequential structure directly and keeps rewards / observation features in float64 libm arithmetic.
#pragma once

#include "uavca_multi.cuh"

namespace uavca {

template <typename P>
struct SeqView {
  P* pos;   // [M][2]
  P* tgt;   // [M][2]
  P* init;  // [M]
  P* prev;  // [M]
};
template <typename P>
__device__ __forceinline__ SeqView<P> seq_view(const StateView& s);
template <>
__device__ __forceinline__ SeqView<float> seq_view<float>(const StateView& s) {
  return SeqView<float>{reinterpret_cast<float*>(s.pos), reinterpret_cast<float*>(s.tgt), s.init, s.prev};
}
template <>
__device__ __forceinline__ SeqView<double> seq_view<double>(const StateView& s) {
  return SeqView<double>{reinterpret_cast<double*>(s.pos64), reinterpret_cast<double*>(s.tgt64), s.init64, s.prev64};
}

// np.linalg.norm of a 2-vector in the position dtype: float32 unfused, float64 with the second product fused (a10)
__device__ __forceinline__ float seq_norm(float dx, float dy) { return n32(dx, dy); }
__device__ __forceinline__ double seq_norm(double dx, double dy) { return sqrt(sq64(dx, dy)); }
__device__ __forceinline__ float seq_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double seq_sub(double a, double b) { return __dsub_rn(a, b); }

// what neighbours are ordered by: the squared distance in the float32 world (as the warp kernels do;
// sqrt is monotone), the distance itself in the float64 world (as the reference's argsort does)
__device__ __forceinline__ float seq_key(float dx, float dy) { return sq32(dx, dy); }
__device__ __forceinline__ double seq_key(double dx, double dy) { return sqrt(sq64(dx, dy)); }
__device__ __forceinline__ float seq_key_to_dist(float k) { return __fsqrt_rn(k); }
__device__ __forceinline__ double seq_key_to_dist(double k) { return k; }

__device__ __forceinline__ double seq_wrap(double x) { return atan2(sin(x), cos(x)); }

// The two nearest other UAVs of UAV i (uav_agent.py:44-64) by distance in the position dtype.  Exact ties: the float32
// world keeps the warp kernels' order (ring offset); the float64 world the reference's own — the
// candidate list is built in agent order and ndarray.argsort is stable below 17 elements.
template <typename P>
__device__ __forceinline__ void seq_nearest2(const P* pos, int N, int i, int& j1, int& j2, P& d1, P& d2) {
  j1 = j2 = -1;
  d1 = d2 = (P)INFINITY;
  const P px = pos[2 * i], py = pos[2 * i + 1];
  for (int k = 1; k < N; ++k) {
    int j;
    if (sizeof(P) == 4) { j = i + k; j -= j >= N ? N : 0; }
    else { j = k - 1; j += j >= i ? 1 : 0; }
    const P d = seq_key(seq_sub(pos[2 * j], px), seq_sub(pos[2 * j + 1], py));
    if (d < d1) { d2 = d1; j2 = j1; d1 = d; j1 = j; }
    else if (d < d2) { d2 = d; j2 = j; }
  }
  d1 = seq_key_to_dist(d1);
  d2 = seq_key_to_dist(d2);
}

// MultiUAVWorld2D._get_obs (multi_uav_world_2d.py:60-109) in float64 libm arithmetic on positions of type P
template <typename P>
__device__ __forceinline__ void seq_obs(const Consts& c, const P* pos, const double* vel, const P* tgt, int N, int i, float* o) {
  const double vx = vel[2 * i], vy = vel[2 * i + 1];
  const double th = atan2(vy, vx);
  const double kInvPi = 0.3183098861837907;
  o[0] = (float)(sqrt(sq64(vx, vy)) * c.inv_vm2);
  o[1] = (float)(th * kInvPi);
  const P tdx = seq_sub(tgt[2 * i], pos[2 * i]), tdy = seq_sub(tgt[2 * i + 1], pos[2 * i + 1]);
  o[2] = (float)((double)seq_norm(tdx, tdy) * c.inv_diag_d);
  o[3] = (float)(seq_wrap(atan2((double)tdy, (double)tdx) - th) * kInvPi);
  int nb[2];
  P dd[2];
  seq_nearest2(pos, N, i, nb[0], nb[1], dd[0], dd[1]);
  bool have = true;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    have = have && nb[k] >= 0 && (sizeof(P) == 4 ? (float)dd[k] < c.dsense : (double)dd[k] < c.dsense_d);
    float* ok = o + 4 + 3 * k;
    if (have) {
      const int j = nb[k];
      ok[0] = sizeof(P) == 4 ? __fdiv_rn((float)dd[k], c.dsense) : (float)((double)dd[k] / c.dsense_d);
      ok[1] = (float)(seq_wrap(atan2((double)seq_sub(pos[2 * j + 1], pos[2 * i + 1]), (double)seq_sub(pos[2 * j], pos[2 * i])) - th) * kInvPi);
      ok[2] = (float)(seq_wrap(atan2(vel[2 * j + 1], vel[2 * j]) - th) * kInvPi);
    } else {
      ok[0] = 1.0f;
      ok[1] = (float)(seq_wrap((3.141592653589793 + th) - th) * kInvPi);
      ok[2] = 0.0f;
    }
  }
}

// MultiUAVWorld2D.reset (multi_uav_world_2d.py:116-168) for one env, sequential rejection sampling (float32 world) or
// the ring of reset(circular=True) (:157-163, float64 world).  The caller has folded the finished episode.
template <typename P>
__device__ __forceinline__ void seq_reset_env(const KernelArgs& a, int b, unsigned episode) {
  const Consts& c = a.c;
  const int N = a.N;
  const SeqView<P> v = seq_view<P>(a.s);
  const size_t m0 = (size_t)b * N;
  const long long env_global = c.env_base + b;
  if (sizeof(P) == 8) {
    for (int i = 0; i < N; ++i) {
      const double4 rg = a.ring64[i];  // host-computed: the reference's math.cos / math.sin are glibc's
      const double px = rg.x, py = rg.y, tx = rg.z, ty = rg.w;
      const size_t m = m0 + i;
      v.pos[2 * m] = (P)px; v.pos[2 * m + 1] = (P)py; v.tgt[2 * m] = (P)tx; v.tgt[2 * m + 1] = (P)ty;
      const double ini = sqrt(sq64(tx - px, ty - py));
      v.init[m] = (P)ini; v.prev[m] = (P)ini;
      a.s.vel[m] = make_double2(0.0, 0.0);
      a.s.flags[m] = 0;
      a.s.pos[m] = make_float2((float)px, (float)py);  // float32 mirrors for readers of the common fields
      a.s.tgt[m] = make_float2((float)tx, (float)ty);
      a.s.init[m] = (float)ini; a.s.prev[m] = (float)ini;
    }
    return;
  }
  if (c.reset_source == UAVCA_SOURCE_POOL && a.pool.pos != nullptr) {
    const long long p = (env_global + (long long)episode) % a.pool_envs;
    for (int i = 0; i < N; ++i) {
      const size_t m = m0 + i, pm = (size_t)p * N + i;
      a.s.pos[m] = a.pool.pos[pm]; a.s.tgt[m] = a.pool.tgt[pm]; a.s.vel[m] = a.pool.vel[pm];
      a.s.init[m] = a.pool.init[pm]; a.s.prev[m] = a.pool.prev[pm]; a.s.flags[m] = a.pool.flags[pm];
    }
    return;
  }
  for (int i = 0; i < N; ++i) {
    a.s.vel[m0 + i] = make_double2(0.0, 0.0);
    a.s.flags[m0 + i] = 0;
  }
  for (int i = 0; i < N; ++i) {  // :126-137
    for (unsigned att = 0;; ++att) {
      const float2 q = draw_pair(c, env_global, episode, kStreamPos, i, att, c.lox, c.hix, c.loy, c.hiy);
      bool rej = false;
      for (int j = 0; j < i && !rej; ++j) {
        const float2 o = a.s.pos[m0 + j];
        rej = n32(__fsub_rn(o.x, q.x), __fsub_rn(o.y, q.y)) <= c.two_r;
      }
      if (!rej || att + 1u >= kMaxResetAttempts) { a.s.pos[m0 + i] = q; break; }
    }
  }
  for (int i = 0; i < N; ++i) {  // :140-155
    const float2 p = a.s.pos[m0 + i];
    for (unsigned att = 0;; ++att) {
      const float2 q = draw_pair(c, env_global, episode, kStreamTgt, i, att, c.lox, c.hix, c.loy, c.hiy);
      bool rej = n32(__fsub_rn(q.x, p.x), __fsub_rn(q.y, p.y)) <= c.two_r;
      for (int j = 0; j < i && !rej; ++j) {
        const float2 o = a.s.tgt[m0 + j];
        rej = n32(__fsub_rn(o.x, q.x), __fsub_rn(o.y, q.y)) <= c.two_r;
      }
      if (!rej || att + 1u >= kMaxResetAttempts) { a.s.tgt[m0 + i] = q; break; }
    }
    const float2 t = a.s.tgt[m0 + i];
    const float ini = n32(__fsub_rn(t.x, p.x), __fsub_rn(t.y, p.y));
    a.s.init[m0 + i] = ini; a.s.prev[m0 + i] = ini;
  }
}

__device__ __forceinline__ void seq_fold_episode(const KernelArgs& a, int b, unsigned episode, int steps) {
  if (episode > 0u) {
    atomicAdd(a.s.stats + 0, 1ull);
    atomicAdd(a.s.stats + 1, (unsigned long long)a.s.reach[b]);
    atomicAdd(a.s.stats + 2, (unsigned long long)a.s.coll[b]);
    atomicAdd(a.s.stats + 3, (unsigned long long)steps);
    if (a.c.track_scores) fold_scores(a.s, b);
  }
  a.s.steps[b] = 0; a.s.reach[b] = 0; a.s.coll[b] = 0;
  a.s.episode[b] = episode + 1u;
}

template <typename P>
__device__ __forceinline__ void seq_obs_env(const KernelArgs& a, int b, float* out) {
  const SeqView<P> v = seq_view<P>(a.s);
  const size_t m0 = (size_t)b * a.N;
  for (int i = 0; i < a.N; ++i)
    seq_obs<P>(a.c, v.pos + 2 * m0, reinterpret_cast<const double*>(a.s.vel + m0), v.tgt + 2 * m0, a.N, i, out + (m0 + i) * 10);
}

// MultiUAVWorld2D.step (multi_uav_world_2d.py:177-241) for env b, UAV after UAV as the reference does it.
template <typename P>
__global__ void __launch_bounds__(kThreads) step_multi_seq_kernel(const __grid_constant__ KernelArgs a) {
  const int b = blockIdx.x * kThreads + threadIdx.x;
  if (b >= a.B) return;
  const Consts& c = a.c;
  const int N = a.N;
  const SeqView<P> v = seq_view<P>(a.s);
  const size_t m0 = (size_t)b * N;
  P* pos = v.pos + 2 * m0;
  const P* tgt = v.tgt + 2 * m0;
  double* vel = reinterpret_cast<double*>(a.s.vel + m0);
  int reach_inc = 0, coll_inc = 0;
  bool done0 = false, all_done = true, any_done = false;
  double live = 0.0, r0 = 0.0;
  for (int i = 0; i < N; ++i) {
    const size_t m = m0 + i;
    unsigned flags = a.s.flags[m];
    const bool parked = (flags & UAVCA_FLAG_PARKED) != 0u;
    P prev_d = (P)0, dist = (P)0;
    if (!parked) {  // UAVAgent.step (uav_agent.py:23-36)
      double ax, ay;  // the action as the reference's float64 arithmetic sees it (uav_agent.py:26)
      if (a.io.action64 != nullptr) {
        const double2 q = a.io.action64[m];
        ax = q.x; ay = q.y;
      } else {
        const float2 act = map_action(a.io.action[m], a.io.action_mode, c);
        ax = (double)act.x; ay = (double)act.y;
      }
      double vx = vel[2 * i], vy = vel[2 * i + 1];
      const double dvx = clipd(__ddiv_rn(__dsub_rn(ax, vx), c.tau), -c.amax, c.amax);
      const double dvy = clipd(__ddiv_rn(__dsub_rn(ay, vy), c.tau), -c.amax, c.amax);
      vx = clipd(__dadd_rn(vx, __dmul_rn(dvx, c.tau)), -c.vmax, c.vmax);
      vy = clipd(__dadd_rn(vy, __dmul_rn(dvy, c.tau)), -c.vmax, c.vmax);
      pos[2 * i] = (P)__dadd_rn((double)pos[2 * i], __dmul_rn(vx, c.tau));  // one rounding in either dtype
      pos[2 * i + 1] = (P)__dadd_rn((double)pos[2 * i + 1], __dmul_rn(vy, c.tau));
      vel[2 * i] = vx; vel[2 * i + 1] = vy;
      prev_d = v.prev[m];
      dist = seq_norm(seq_sub(tgt[2 * i], pos[2 * i]), seq_sub(tgt[2 * i + 1], pos[2 * i + 1]));
    }
    const P px = pos[2 * i], py = pos[2 * i + 1];
    const double dth = seq_wrap(atan2((double)seq_sub(tgt[2 * i + 1], py), (double)seq_sub(tgt[2 * i], px)) - atan2(vel[2 * i + 1], vel[2 * i]));
    // reward (:188-195) in the reference's dtype sequence
    const P ini = v.init[m];
    double mm = c.vm2 / (double)ini;
    if (1.0 < mm) mm = 1.0;
    double r = 0.0 - 0.01 * mm;
    double f;
    if (sizeof(P) == 4) {
      r += 50.0 * ((double)__fsub_rn((float)prev_d, (float)dist) / c.vm2);
      const float q = __fdiv_rn((float)dist, __fmul_rn(1.5f, (float)ini));
      f = (double)((r > 0) ? __fsub_rn(1.0f, q) : __fadd_rn(1.0f, q));
    } else {
      r += 50.0 * (((double)prev_d - (double)dist) / c.vm2);
      const double q = (double)dist / (1.5 * (double)ini);
      f = (r > 0) ? (1.0 - q) : (1.0 + q);
    }
    r *= f;
    r -= 0.01 * fabs(dth);
    // nearest in-range neighbour on the mixed old/new positions (:198-210): UAVs j < i have moved already
    P dmin = (P)INFINITY;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const P d = seq_norm(seq_sub(pos[2 * j], px), seq_sub(pos[2 * j + 1], py));
      dmin = d < dmin ? d : dmin;
    }
    bool collision = false;
    const bool sensed = sizeof(P) == 4 ? (float)dmin < c.dsense : (double)dmin < c.dsense_d;
    if (sensed) {
      const bool soft = sizeof(P) == 4 ? (float)dmin <= c.two_r : (double)dmin <= c.two_r_d;
      const bool hard = sizeof(P) == 4 ? (float)dmin <= c.two_h : (double)dmin <= c.two_h_d;
      if (soft) { r = -2.0; collision = true; }
      if (hard && !parked && !(flags & UAVCA_FLAG_COLLIDED)) { coll_inc += 1; flags |= UAVCA_FLAG_COLLIDED; }
    }
    const double vsq = sq64(vel[2 * i], vel[2 * i + 1]);
    const bool inside = (double)px >= c.lox && (double)px <= c.hix && (double)py >= c.loy && (double)py <= c.hiy;
    const bool close = sizeof(P) == 4 ? (float)dist < c.reach_dist : (double)dist < c.reach_dist_d;
    bool d;
    if (close && !collision && vsq < c.reach_speed_sq) {  // :218-223
      d = true;
      if (!parked) reach_inc += 1;
      flags |= UAVCA_FLAG_PARKED;
      const double2 fv = finish_velocity(vel[2 * i], vel[2 * i + 1], vsq);
      vel[2 * i] = fv.x; vel[2 * i + 1] = fv.y;
      r += 10.0;
    } else if (!inside) {
      d = a.io.evaluate == 0;  // :224-225
    } else {
      d = false;
    }
    v.prev[m] = dist;  // :229
    a.s.flags[m] = (uint8_t)flags;
    a.io.reward[m] = (float)r;
    a.io.done[m] = (uint8_t)d;
    if (i == 0) { done0 = d; r0 = r; }
    all_done &= d; any_done |= d;
    live += d ? 0.0 : (double)(float)r;
    if (!(fabs(r) <= 3.4e38) | !(fabs((double)px) + fabs((double)py) <= 3.4e38)) atomicAdd(a.s.stats + 6, 1ull);
  }
  if (sizeof(P) == 8)  // float32 mirrors of the float64 state
    for (int i = 0; i < N; ++i) {
      a.s.pos[m0 + i] = make_float2((float)pos[2 * i], (float)pos[2 * i + 1]);
      a.s.prev[m0 + i] = (float)v.prev[m0 + i];
    }
  seq_obs_env<P>(a, b, a.io.obs);  // :233-235
  if (a.io.final_obs)
    for (int k = 0; k < N * 10; ++k) a.io.final_obs[m0 * 10 + k] = a.io.obs[m0 * 10 + k];
  const int steps_new = a.s.steps[b] + 1;  // :238
  if (reach_inc) a.s.reach[b] += reach_inc;
  if (coll_inc) a.s.coll[b] += coll_inc;
  if (c.track_scores) {
    double2 sc = a.s.score[b];
    sc.x += (double)(float)r0; sc.y += live;
    a.s.score[b] = sc;
  }
  const bool rs = ((c.reset_mode & UAVCA_RESET_ON_DONE0) && done0) || ((c.reset_mode & UAVCA_RESET_ON_ALL_DONE) && all_done) ||
                  ((c.reset_mode & UAVCA_RESET_ON_ANY_DONE) && any_done) || steps_new >= c.steps_limit;
  if (a.io.reset_mask) a.io.reset_mask[b] = (uint8_t)rs;
  a.s.steps[b] = steps_new;
  if (rs) {
    const unsigned episode = a.s.episode[b];
    seq_fold_episode(a, b, episode, steps_new);
    seq_reset_env<P>(a, b, episode);
    seq_obs_env<P>(a, b, a.io.obs);
  }
}

template <typename P>
__global__ void __launch_bounds__(kThreads) reset_multi_seq_kernel(const __grid_constant__ KernelArgs a, const uint8_t* mask) {
  const int b = blockIdx.x * kThreads + threadIdx.x;
  if (b >= a.B) return;
  if (mask != nullptr && mask[b] == 0) return;
  const unsigned episode = a.s.episode[b];
  seq_fold_episode(a, b, episode, a.s.steps[b]);
  seq_reset_env<P>(a, b, episode);
  if (a.io.obs) seq_obs_env<P>(a, b, a.io.obs);
}

template <typename P>
__global__ void __launch_bounds__(kThreads) observe_multi_seq_kernel(const __grid_constant__ KernelArgs a) {
  const int b = blockIdx.x * kThreads + threadIdx.x;
  if (b >= a.B) return;
  seq_obs_env<P>(a, b, a.io.obs);
}

}  // namespace uavca
