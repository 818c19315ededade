/*
 * uav_oracle.c — CPU restatement of the reference environment step.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call it.  The product package never does.
 *
 * Parity pinning: the reference ships no golden vectors or tests.  This restatement is
 * pinned by executing the unmodified reference in the build container (oracle/ref_loader.py) and comparing
 * bit-for-bit: oracle/gen_golden.py writes the tests/golden npz files from the literal reference and
 * tests/test_oracle_golden.py replays them through this file (flags, float32 positions, float64 velocities,
 * float64 rewards and float64 observations all compared for exact equality).
 *
 * Each function cites the reference lines it restates (paths relative to the reference checkout,
 * gym_uav_collision_avoidance/envs/...).  Semantics, operation order and dtypes follow the reference as it
 * executes under NumPy 2 (NEP 50 weak Python scalars); the code shape does not.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -pthread (see oracle/Makefile).  -ffp-contract=off matters:
 * every fused multiply-add below is an explicit fma() where the reference's BLAS has one.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/uavca.h"

#include <pthread.h>
#include <unistd.h>

#define PI_D 3.141592653589793 /* math.pi */

typedef struct uavo_state {
  float* pos;        /* [B][N][2] */
  double* vel;       /* [B][N][2] */
  float* tgt;        /* [B][N][2] */
  float* init;       /* [B][N] */
  float* prev;       /* [B][N] */
  uint8_t* flags;    /* [B][N] */
  int32_t* steps;    /* [B] */
  int32_t* reach;    /* [B] */
  int32_t* coll;     /* [B] */
  uint32_t* episode; /* [B] */
  double* score;     /* [B][2] running scores of the episode in flight (config.track_scores) */
  /* float64 world (config.circular): after reset(circular=True) the reference holds locations and targets as float64
   * arrays (multi_uav_world_2d.py:157-163), so every distance, threshold test and reward term of such an episode is
   * float64.  These fields are the state of that mode; pos/tgt/init/prev above then hold rounded mirrors. */
  double* pos64;     /* [B][N][2] */
  double* tgt64;     /* [B][N][2] */
  double* init64;    /* [B][N] */
  double* prev64;    /* [B][N] */
  uint64_t* stats;   /* [8]: episodes, reach, collisions, steps; double score sums in [4], [5]; non-finite count in [6] */
} uavo_state;

/* stats[4], stats[5] hold doubles; the auto-reset pass runs on the thread pool, so they are added with a CAS loop */
static void atomic_add_double(uint64_t* slot, double v) {
  uint64_t old = __atomic_load_n(slot, __ATOMIC_RELAXED), neu;
  do {
    double d;
    memcpy(&d, &old, 8);
    d += v;
    memcpy(&neu, &d, 8);
  } while (!__atomic_compare_exchange_n(slot, &old, neu, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
static void fold_scores(const uavca_config* c, uavo_state* s, int b) {
  if (!c->track_scores || !s->score || !s->stats) return;
  atomic_add_double(&s->stats[4], s->score[2 * b]);
  atomic_add_double(&s->stats[5], s->score[2 * b + 1]);
  s->score[2 * b] = 0.0; s->score[2 * b + 1] = 0.0;
}

/* ---- numpy primitives as they execute in the reference ---------------------------------------------- */

/* np.linalg.norm of a float32 2-vector: sqrtf(x*x + y*y), products and sum rounded separately
 * (OpenBLAS sdot on 2 elements + sqrt). */
static inline float n32(float dx, float dy) {
  float a = dx * dx;
  float b = dy * dy;
  float s = a + b;
  return sqrtf(s);
}
static inline float s32(float dx, float dy) {
  float a = dx * dx;
  float b = dy * dy;
  return a + b;
}

/* np.linalg.norm of a float64 2-vector: the ddot kernel fuses the second product. */
static inline double n64(double x, double y) { return sqrt(fma(y, y, x * x)); }

static inline double clipd(double v, double lo, double hi) {
  /* np.clip == minimum(maximum(v, lo), hi); NaN propagates */
  if (v != v) return v;
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}

static inline double wrap(double x) { return atan2(sin(x), cos(x)); }

/* ---- Philox4x32-10 (Salmon et al., SC'11), the on-device reset stream ------------------------------- */

static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                 uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53-bit uniform in [0,1) from two words (same construction as numpy's random_sample). */
static inline double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

enum { STREAM_POS = 0, STREAM_TGT = 1, STREAM_VEL = 2, STREAM_ACT = 3 };

/* One draw of np.random.uniform(low, high, size=(2,)).astype(float32): low + (high-low)*u per component. */
static inline void draw_pair(const uavca_config* c, int64_t env_global, uint32_t episode, int stream, int uav,
                             uint32_t attempt, double lo_x, double hi_x, double lo_y, double hi_y, float* x, float* y) {
  uint32_t r[4];
  philox4x32_10((uint32_t)env_global, episode, ((uint32_t)stream << 16) | (uint32_t)uav, attempt,
                (uint32_t)c->seed, (uint32_t)(c->seed >> 32), r);
  double ux = u53(r[0], r[1]);
  double uy = u53(r[2], r[3]);
  double rx = (hi_x - lo_x) * ux;
  double ry = (hi_y - lo_y) * uy;
  *x = (float)(lo_x + rx);
  *y = (float)(lo_y + ry);
}

#define MAX_RESET_ATTEMPTS 4096u

/* ---- observation: MultiUAVWorld2D._get_obs, multi_uav_world_2d.py:60-109 ---------------------------- */

static void obs_multi(const uavca_config* c, const float* pos, const double* vel, const float* tgt, int N, int i,
                      double* o) {
  const double vm2 = n64(c->max_speed, c->max_speed);      /* np.linalg.norm(agent.max_speed)  :62 */
  const double diag = n64(c->x_size, c->y_size);           /* self.map_diagonal_size           :17 */
  const float dsense32 = (float)c->d_sense;                /* python scalar, weak -> float32   :77 */
  const double vx = vel[2 * i], vy = vel[2 * i + 1];
  const float px = pos[2 * i], py = pos[2 * i + 1];
  const double th = atan2(vy, vx);                          /* :63 */
  o[0] = n64(vx, vy) / vm2;                                 /* :62 */
  o[1] = th / PI_D;                                         /* :64 */
  const float tdx = tgt[2 * i] - px, tdy = tgt[2 * i + 1] - py;
  o[2] = (double)n32(tdx, tdy) / diag;                      /* :67-68 */
  o[3] = wrap(atan2((double)tdy, (double)tdx) - th) / PI_D; /* :69-72 */

  /* uavs_in_range (uav_agent.py:44-64): in-range neighbours ascending by float32 distance.  Exact-distance
   * ties have no defined order in the reference (unstable argsort); we refine the order by (squared
   * distance, ring offset (j - i) mod N), which agrees with the reference whenever it is defined. */
  int j1 = -1, j2 = -1;
  float s1 = INFINITY, s2 = INFINITY;
  for (int k = 1; k < N; ++k) {
    int j = i + k; if (j >= N) j -= N;
    float s = s32(pos[2 * j] - px, pos[2 * j + 1] - py);
    if (s < s1) { s2 = s1; j2 = j1; s1 = s; j1 = j; }
    else if (s < s2) { s2 = s; j2 = j; }
  }
  const int nb[2] = {j1, j2};
  const float sq[2] = {s1, s2};
  int have = 1;
  for (int k = 0; k < 2; ++k) {
    double* ok = o + 4 + 3 * k;
    float d = 0.f;
    if (have && nb[k] >= 0) { d = sqrtf(sq[k]); have = d < dsense32; } else have = 0;
    if (have) {
      int j = nb[k];
      float dx = pos[2 * j] - px, dy = pos[2 * j + 1] - py;
      ok[0] = (double)(d / dsense32);                                    /* :77 float32 quotient */
      ok[1] = wrap(atan2((double)dy, (double)dx) - th) / PI_D;           /* :78-81 */
      ok[2] = wrap(atan2(vel[2 * j + 1], vel[2 * j]) - th) / PI_D;       /* :82-85 */
    } else {
      ok[0] = 1.0;                                                        /* :77 */
      ok[1] = wrap((PI_D + th) - th) / PI_D;                              /* :78-81 (+1, 0.99.. or -1) */
      ok[2] = wrap(th - th) / PI_D;                                       /* :82-85 */
    }
  }
}

/* ---- observation: UAVWorld2D._get_obs, uav_world_2d.py:77-112 --------------------------------------- */

static void obs_single(const uavca_config* c, const float* pos, const double* vel, const float* tgt, int vel_is_f32,
                       double* o) {
  const double diag = n64(c->x_size, c->y_size);             /* :17 */
  const double vx = vel[0], vy = vel[1];
  /* right after reset() the speed is a float32 array (:122) and its norm is a float32 norm */
  const double speed = vel_is_f32 ? (double)n32((float)vx, (float)vy) : n64(vx, vy);
  const double th = atan2(vy, vx);                           /* :89 */
  const float tdx = tgt[0] - pos[0], tdy = tgt[1] - pos[1];
  o[0] = speed / c->max_speed;                               /* :88 */
  o[1] = th / PI_D;                                          /* :90 */
  o[2] = (double)n32(tdx, tdy) / diag;                       /* :96-97 */
  o[3] = wrap(atan2((double)tdy, (double)tdx) - th) / PI_D;  /* :91-94 */
}

/* ---- action mapping done by the callers: test_sac_multi.py:77-80, test_pytorch_multi.py:80 ---------- */

static inline void map_action(const uavca_config* c, int mode, float a0, float a1, float* ox, float* oy) {
  if (mode == UAVCA_ACTION_POLAR) {
    float v = (a0 / 2.0f + 0.5f) * (float)c->polar_scale;
    float th = a1 * (float)PI_D;
    *ox = v * (float)cos((double)th);
    *oy = v * (float)sin((double)th);
  } else if (mode == UAVCA_ACTION_SCALED) {
    *ox = a0 * (float)c->max_speed;
    *oy = a1 * (float)c->max_speed;
  } else {
    *ox = a0; *oy = a1;
  }
}

/* ---- reset: MultiUAVWorld2D.reset, multi_uav_world_2d.py:116-175 ------------------------------------ */

static void reset_env_multi(const uavca_config* c, uavo_state* s, const uavo_state* pool, int pool_envs, int b) {
  const int N = c->num_agents;
  float* pos = s->pos + (size_t)b * N * 2;
  double* vel = s->vel + (size_t)b * N * 2;
  float* tgt = s->tgt + (size_t)b * N * 2;
  float* init = s->init + (size_t)b * N;
  float* prev = s->prev + (size_t)b * N;
  uint8_t* flags = s->flags + (size_t)b * N;
  const int64_t env_global = c->env_index_base + b;
  const uint32_t ep = s->episode[b];

  /* fold the finished episode into the running totals */
  if (s->stats && ep > 0) { /* atomic: the auto-reset pass runs on the thread pool */
    __atomic_fetch_add(&s->stats[0], 1, __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[1], (uint64_t)s->reach[b], __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[2], (uint64_t)s->coll[b], __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[3], (uint64_t)s->steps[b], __ATOMIC_RELAXED);
    fold_scores(c, s, b);
  }

  if (c->reset_source == UAVCA_SOURCE_POOL && pool && pool_envs > 0) {
    const size_t p = (size_t)((env_global + (int64_t)ep) % pool_envs);
    memcpy(pos, pool->pos + p * N * 2, sizeof(float) * N * 2);
    memcpy(vel, pool->vel + p * N * 2, sizeof(double) * N * 2);
    memcpy(tgt, pool->tgt + p * N * 2, sizeof(float) * N * 2);
    memcpy(init, pool->init + p * N, sizeof(float) * N);
    memcpy(prev, pool->prev + p * N, sizeof(float) * N);
    memcpy(flags, pool->flags + p * N, N);
  } else {
    const double lox = -c->x_size / 2.0, hix = c->x_size / 2.0; /* :19-20 */
    const double loy = -c->y_size / 2.0, hiy = c->y_size / 2.0;
    const float two_r = (float)(2.0 * c->collider_radius);      /* float32 norm <= python float: weak */
    for (int i = 0; i < N; ++i) {                               /* :118-123 */
      vel[2 * i] = 0.0; vel[2 * i + 1] = 0.0; flags[i] = 0;
    }
    for (int i = 0; i < N; ++i) {                               /* :126-137 */
      for (uint32_t a = 0;; ++a) {
        float x, y;
        draw_pair(c, env_global, ep, STREAM_POS, i, a, lox, hix, loy, hiy, &x, &y);
        int rej = 0;
        for (int j = 0; j < i && !rej; ++j) rej = n32(pos[2 * j] - x, pos[2 * j + 1] - y) <= two_r;
        if (!rej || a + 1 >= MAX_RESET_ATTEMPTS) { pos[2 * i] = x; pos[2 * i + 1] = y; break; }
      }
    }
    for (int i = 0; i < N; ++i) {                               /* :140-155 */
      for (uint32_t a = 0;; ++a) {
        float x, y;
        draw_pair(c, env_global, ep, STREAM_TGT, i, a, lox, hix, loy, hiy, &x, &y);
        int rej = n32(x - pos[2 * i], y - pos[2 * i + 1]) <= two_r;
        for (int j = 0; j < i && !rej; ++j) rej = n32(tgt[2 * j] - x, tgt[2 * j + 1] - y) <= two_r;
        if (!rej || a + 1 >= MAX_RESET_ATTEMPTS) { tgt[2 * i] = x; tgt[2 * i + 1] = y; break; }
      }
    }
    for (int i = 0; i < N; ++i) {
      init[i] = n32(tgt[2 * i] - pos[2 * i], tgt[2 * i + 1] - pos[2 * i + 1]); /* :154 */
      prev[i] = init[i];                                                        /* :155 */
    }
  }
  s->steps[b] = 0; s->reach[b] = 0; s->coll[b] = 0; /* :166-168 */
  s->episode[b] = ep + 1;
}

/* ---- reset: UAVWorld2D.reset, uav_world_2d.py:119-135 ----------------------------------------------- */

static void reset_env_single(const uavca_config* c, uavo_state* s, const uavo_state* pool, int pool_envs, int b) {
  float* pos = s->pos + (size_t)b * 2;
  double* vel = s->vel + (size_t)b * 2;
  float* tgt = s->tgt + (size_t)b * 2;
  const int64_t env_global = c->env_index_base + b;
  const uint32_t ep = s->episode[b];
  if (s->stats && ep > 0) { /* atomic: the auto-reset pass runs on the thread pool */
    __atomic_fetch_add(&s->stats[0], 1, __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[1], (uint64_t)s->reach[b], __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[2], (uint64_t)s->coll[b], __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[3], (uint64_t)s->steps[b], __ATOMIC_RELAXED);
    fold_scores(c, s, b);
  }
  if (c->reset_source == UAVCA_SOURCE_POOL && pool && pool_envs > 0) {
    const size_t p = (size_t)((env_global + (int64_t)ep) % pool_envs);
    memcpy(pos, pool->pos + p * 2, sizeof(float) * 2);
    memcpy(vel, pool->vel + p * 2, sizeof(double) * 2);
    memcpy(tgt, pool->tgt + p * 2, sizeof(float) * 2);
    s->init[b] = pool->init[p];
    s->prev[b] = pool->prev[p];
    s->flags[b] = pool->flags[p];
  } else {
    const double lox = -c->x_size / 2.0, hix = c->x_size / 2.0;
    const double loy = -c->y_size / 2.0, hiy = c->y_size / 2.0;
    float x, y;
    draw_pair(c, env_global, ep, STREAM_POS, 0, 0, lox, hix, loy, hiy, &x, &y);           /* :121 */
    pos[0] = x; pos[1] = y;
    draw_pair(c, env_global, ep, STREAM_VEL, 0, 0, -c->max_speed, c->max_speed, -c->max_speed, c->max_speed, &x, &y); /* :122 */
    vel[0] = (double)x; vel[1] = (double)y;
    draw_pair(c, env_global, ep, STREAM_TGT, 0, 0, lox, hix, loy, hiy, &x, &y);           /* :126 */
    tgt[0] = x; tgt[1] = y;
    s->init[b] = n32(tgt[0] - pos[0], tgt[1] - pos[1]);                                    /* :129 */
    s->prev[b] = s->init[b];                                                               /* :130 */
    s->flags[b] = 0;
  }
  s->steps[b] = 0; s->reach[b] = 0; s->coll[b] = 0;                                        /* :131 */
  s->episode[b] = ep + 1;
}

/* ---- step: MultiUAVWorld2D.step (multi_uav_world_2d.py:177-241) + UAVAgent (uav_agent.py:23-64) ----- */

static void step_env_multi(const uavca_config* c, uavo_state* s, int b, const float* action, const double* a64, int action_mode,
                           int evaluate, double* obs, double* reward, uint8_t* done) {
  const int N = c->num_agents;
  float* pos = s->pos + (size_t)b * N * 2;
  double* vel = s->vel + (size_t)b * N * 2;
  const float* tgt = s->tgt + (size_t)b * N * 2;
  const float* init = s->init + (size_t)b * N;
  float* prev = s->prev + (size_t)b * N;
  uint8_t* flags = s->flags + (size_t)b * N;
  const double tau = c->tau, amax = c->max_acceleration, vmax = c->max_speed;
  const double vm2 = n64(vmax, vmax);                         /* np.linalg.norm(agent.max_speed) :183 */
  const float two_r = (float)(2.0 * c->collider_radius);      /* :203 weak python float vs float32 */
  const float two_h = (float)(2.0 * c->hard_collision_radius);/* :207 */
  const float dsense32 = (float)c->d_sense;                   /* uav_agent.py:52 */
  const float reach32 = (float)c->reach_distance;             /* :218 */
  const double lox = -c->x_size / 2.0, hix = c->x_size / 2.0, loy = -c->y_size / 2.0, hiy = c->y_size / 2.0;

  for (int i = 0; i < N; ++i) { /* sequential: UAV i sees j<i moved, j>i not yet moved (:181) */
    const int parked = flags[i] & UAVCA_FLAG_PARKED;
    float prev_d, dist;
    if (parked) {                                             /* uav_agent.py:24-25 returns ints 0, 0 */
      prev_d = 0.f; dist = 0.f;
    } else {
      double ax, ay; /* the action as float64 arithmetic sees it: a float32 action widens exactly, a float64 one (a64: what
                        the reference's own loops build, test_sac_multi.py:77-80) enters as it is */
      if (a64) { ax = a64[2 * i]; ay = a64[2 * i + 1]; }
      else { float fx, fy; map_action(c, action_mode, action[2 * i], action[2 * i + 1], &fx, &fy); ax = (double)fx; ay = (double)fy; }
      double dvx = clipd((ax - vel[2 * i]) / tau, -amax, amax);       /* uav_agent.py:26 */
      double dvy = clipd((ay - vel[2 * i + 1]) / tau, -amax, amax);
      double vx = clipd(vel[2 * i] + dvx * tau, -vmax, vmax);                 /* :27 */
      double vy = clipd(vel[2 * i + 1] + dvy * tau, -vmax, vmax);
      pos[2 * i] = (float)((double)pos[2 * i] + vx * tau);                    /* :28-29 float32 += float64 */
      pos[2 * i + 1] = (float)((double)pos[2 * i + 1] + vy * tau);
      vel[2 * i] = vx; vel[2 * i + 1] = vy;                                   /* :30 */
      prev_d = prev[i];                                                       /* :32 */
      dist = n32(tgt[2 * i] - pos[2 * i], tgt[2 * i + 1] - pos[2 * i + 1]);   /* :33 */
    }
    const float px = pos[2 * i], py = pos[2 * i + 1];
    const float tdx = tgt[2 * i] - px, tdy = tgt[2 * i + 1] - py;
    double dth = atan2((double)tdy, (double)tdx) - atan2(vel[2 * i + 1], vel[2 * i]); /* :184-185 */
    dth = wrap(dth);                                                                   /* :186 */

    double m = vm2 / (double)init[i];                          /* :189 float64 / float32 */
    if (1.0 < m) m = 1.0;                                      /* python min(x, 1) */
    double r = 0.0 - 0.01 * m;
    r += 50.0 * ((double)(prev_d - dist) / vm2);               /* :190 float32 subtraction first */
    {
      float q = dist / (1.5f * init[i]);                       /* :192/:194 float32 product and quotient */
      float f = (r > 0) ? (1.0f - q) : (1.0f + q);
      r *= (double)f;
    }
    r -= 0.01 * fabs(dth);                                     /* :195 */

    /* nearest in-range neighbour on the mixed old/new positions (:198-210; testing the two nearest of an
     * ascending list is the same as testing the minimum) */
    float smin = INFINITY;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      float sj = s32(pos[2 * j] - px, pos[2 * j + 1] - py);
      if (sj < smin) smin = sj;
    }
    int collision = 0;
    if (smin < INFINITY) {
      float dmin = sqrtf(smin);
      if (dmin < dsense32) {
        if (dmin <= two_r) { r = -2.0; collision = 1; }        /* :203-205 */
        if (dmin <= two_h && !parked && !(flags[i] & UAVCA_FLAG_COLLIDED)) { /* :207-210 */
          s->coll[b] += 1;
          flags[i] |= UAVCA_FLAG_COLLIDED;
        }
      }
    }
    const double speed = n64(vel[2 * i], vel[2 * i + 1]);      /* :214 */
    const int oob = !((double)px >= lox && (double)px <= hix && (double)py >= loy && (double)py <= hiy); /* :213,224 */
    int d;
    if (dist < reach32 && !collision && speed < c->reach_speed) { /* :218 */
      d = 1;
      if (!parked) s->reach[b] += 1;                           /* :220-221 */
      flags[i] |= UAVCA_FLAG_PARKED;                           /* finish(): uav_agent.py:38-42 */
      double nv = n64(vel[2 * i], vel[2 * i + 1]);
      double fx = vel[2 * i] / nv * 0.001, fy = vel[2 * i + 1] / nv * 0.001;
      if (fx != fx || fy != fy) { fx = 0.0; fy = 0.0; }
      vel[2 * i] = fx; vel[2 * i + 1] = fy;
      r += 10.0;                                               /* :223 */
    } else if (oob) {
      d = evaluate ? 0 : 1;                                    /* :225 */
    } else {
      d = 0;
    }
    prev[i] = dist;                                            /* :229 */
    reward[i] = r;
    done[i] = (uint8_t)d;
  }
  for (int i = 0; i < N; ++i) obs_multi(c, pos, vel, tgt, N, i, obs + 10 * i); /* :233-235 */
  s->steps[b] += 1;                                            /* :238 */
  if (c->track_scores && s->score) { /* what the callers accumulate: test_sac_multi.py:105 and :152-156 */
    double live = 0.0;
    for (int i = 0; i < N; ++i) live += reward[i] * (1 - done[i]);
    s->score[2 * b] += reward[0];
    s->score[2 * b + 1] += live;
  }
  if (s->stats)
    for (int i = 0; i < N; ++i)
      if (!isfinite(reward[i]) || !isfinite(pos[2 * i]) || !isfinite(pos[2 * i + 1])) __atomic_fetch_add(&s->stats[6], 1, __ATOMIC_RELAXED);
}

/* ---- step: UAVWorld2D.step, uav_world_2d.py:137-173 ------------------------------------------------- */

static void step_env_single(const uavca_config* c, uavo_state* s, int b, const float* action, const double* a64, int action_mode,
                            double* obs, double* reward, uint8_t* done, float* distance) {
  float* pos = s->pos + (size_t)b * 2;
  double* vel = s->vel + (size_t)b * 2;
  const float* tgt = s->tgt + (size_t)b * 2;
  const double tau = c->tau, amax = c->max_acceleration, vmax = c->max_speed;
  const double lox = -c->x_size / 2.0, hix = c->x_size / 2.0, loy = -c->y_size / 2.0, hiy = c->y_size / 2.0;
  float ax = 0.f, ay = 0.f;
  if (!a64) map_action(c, action_mode, action[0], action[1], &ax, &ay);
  double qx, qy;
  if (a64) { /* a float64 action: float64 arithmetic from the first step on (uav_world_2d.py:142) */
    qx = (a64[0] - vel[0]) / tau;
    qy = (a64[1] - vel[1]) / tau;
  } else if (c->single_f32_first_step && s->steps[b] == 0) {
    /* float32 action minus the float32 reset speed, divided by the weak python float tau: all float32 */
    qx = (double)((ax - (float)vel[0]) / (float)tau);
    qy = (double)((ay - (float)vel[1]) / (float)tau);
  } else {
    qx = ((double)ax - vel[0]) / tau;
    qy = ((double)ay - vel[1]) / tau;
  }
  double dvx = clipd(qx, -amax, amax), dvy = clipd(qy, -amax, amax);   /* :142 */
  double vx = clipd(vel[0] + dvx * tau, -vmax, vmax);                  /* :144 */
  double vy = clipd(vel[1] + dvy * tau, -vmax, vmax);
  pos[0] = (float)((double)pos[0] + vx * tau);                         /* :145-146 */
  pos[1] = (float)((double)pos[1] + vy * tau);
  vel[0] = vx; vel[1] = vy;                                            /* :147 */
  const float tdx = tgt[0] - pos[0], tdy = tgt[1] - pos[1];
  const float dist = n32(tdx, tdy);                                    /* :150 */
  float r = 0.0f - 1.0f / s->init[b];                                  /* :152-153 float32 */
  r = r + 10.0f * (s->prev[b] - dist);                                 /* :154 float32 */
  double dth = wrap(atan2((double)tdy, (double)tdx) - atan2(vy, vx));  /* :155-156 */
  r = r - (float)(0.1 * fabs(dth));                                    /* :157 python float is weak -> float32 */
  const int oob = !((double)pos[0] >= lox && (double)pos[0] <= hix && (double)pos[1] >= loy && (double)pos[1] <= hiy);
  int d;
  if (dist < (float)c->reach_distance) { d = 1; r = r + 1000.0f; s->reach[b] += 1; } /* :159-161 (+ our success counter) */
  else if (oob) d = 1;                                                 /* :162-163 */
  else d = 0;
  obs_single(c, pos, vel, tgt, 0, obs);                                /* :168 */
  if (distance) *distance = dist;                                      /* :169 info["distance"] */
  s->steps[b] += 1;                                                    /* :170 */
  s->prev[b] = dist;                                                   /* :172 */
  *reward = (double)r;
  *done = (uint8_t)d;
  if (c->track_scores && s->score) {
    s->score[2 * b] += (double)r;
    s->score[2 * b + 1] += d ? 0.0 : (double)r;
  }
  if (s->stats && (!isfinite(r) || !isfinite(pos[0]) || !isfinite(pos[1]))) __atomic_fetch_add(&s->stats[6], 1, __ATOMIC_RELAXED);
}

/* ==== float64 world: episodes started by reset(circular=True) ==========================================
 * multi_uav_world_2d.py:157-163 assigns `20 * np.ones(2) * np.array([cos, sin])` — float64 arrays — to location and
 * target_location, so UAVAgent.step's `self.location += dx` (uav_agent.py:28-29), every np.linalg.norm and every
 * comparison of the episode run in float64.  Same control flow as the float32 functions above, restated with the
 * float64 dtype sequence.  Exact-distance ties (the ring is symmetric) keep the candidate-list order of
 * uavs_in_range, i.e. ascending agent index (ndarray.argsort is an insertion sort, hence stable, below 17 elements). */

static void nearest2_f64(const double* pos, int N, int i, int* j1, int* j2, double* d1, double* d2) {
  *j1 = *j2 = -1; *d1 = *d2 = INFINITY;
  for (int j = 0; j < N; ++j) {                               /* uav_agent.py:47-55, candidate order = agent index */
    if (j == i) continue;
    double d = n64(pos[2 * j] - pos[2 * i], pos[2 * j + 1] - pos[2 * i + 1]);
    if (d < *d1) { *d2 = *d1; *j2 = *j1; *d1 = d; *j1 = j; }
    else if (d < *d2) { *d2 = d; *j2 = j; }
  }
}

static void obs_multi_f64(const uavca_config* c, const double* pos, const double* vel, const double* tgt, int N, int i,
                          double* o) {
  const double vm2 = n64(c->max_speed, c->max_speed), diag = n64(c->x_size, c->y_size);
  const double vx = vel[2 * i], vy = vel[2 * i + 1];
  const double th = atan2(vy, vx);
  o[0] = n64(vx, vy) / vm2;
  o[1] = th / PI_D;
  const double tdx = tgt[2 * i] - pos[2 * i], tdy = tgt[2 * i + 1] - pos[2 * i + 1];
  o[2] = n64(tdx, tdy) / diag;
  o[3] = wrap(atan2(tdy, tdx) - th) / PI_D;
  int nb[2]; double dd[2];
  nearest2_f64(pos, N, i, &nb[0], &nb[1], &dd[0], &dd[1]);
  int have = 1;
  for (int k = 0; k < 2; ++k) {
    double* ok = o + 4 + 3 * k;
    have = have && nb[k] >= 0 && dd[k] < c->d_sense;           /* uav_agent.py:52 */
    if (have) {
      int j = nb[k];
      ok[0] = dd[k] / c->d_sense;
      ok[1] = wrap(atan2(pos[2 * j + 1] - pos[2 * i + 1], pos[2 * j] - pos[2 * i]) - th) / PI_D;
      ok[2] = wrap(atan2(vel[2 * j + 1], vel[2 * j]) - th) / PI_D;
    } else {
      ok[0] = 1.0;
      ok[1] = wrap((PI_D + th) - th) / PI_D;
      ok[2] = wrap(th - th) / PI_D;
    }
  }
}

static void mirror_f32(uavo_state* s, int b, int N) {
  for (int i = 0; i < N; ++i) {
    size_t m = (size_t)b * N + i;
    s->pos[2 * m] = (float)s->pos64[2 * m]; s->pos[2 * m + 1] = (float)s->pos64[2 * m + 1];
    s->tgt[2 * m] = (float)s->tgt64[2 * m]; s->tgt[2 * m + 1] = (float)s->tgt64[2 * m + 1];
    s->init[m] = (float)s->init64[m]; s->prev[m] = (float)s->prev64[m];
  }
}

static void reset_env_circular(const uavca_config* c, uavo_state* s, int b) {
  const int N = c->num_agents;
  const uint32_t ep = s->episode[b];
  if (s->stats && ep > 0) {
    __atomic_fetch_add(&s->stats[0], 1, __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[1], (uint64_t)s->reach[b], __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[2], (uint64_t)s->coll[b], __ATOMIC_RELAXED);
    __atomic_fetch_add(&s->stats[3], (uint64_t)s->steps[b], __ATOMIC_RELAXED);
    fold_scores(c, s, b);
  }
  for (int i = 0; i < N; ++i) {                                /* :118-123, :157-163 */
    size_t m = (size_t)b * N + i;
    double theta = 2 * i * PI_D / N;
    s->vel[2 * m] = 0.0; s->vel[2 * m + 1] = 0.0; s->flags[m] = 0;
    s->pos64[2 * m] = 20.0 * cos(theta); s->pos64[2 * m + 1] = 20.0 * sin(theta);
    s->tgt64[2 * m] = 23.0 * cos(theta + PI_D); s->tgt64[2 * m + 1] = 23.0 * sin(theta + PI_D);
    s->init64[m] = n64(s->tgt64[2 * m] - s->pos64[2 * m], s->tgt64[2 * m + 1] - s->pos64[2 * m + 1]);
    s->prev64[m] = s->init64[m];
  }
  mirror_f32(s, b, N);
  s->steps[b] = 0; s->reach[b] = 0; s->coll[b] = 0;
  s->episode[b] = ep + 1;
}

static void step_env_multi_f64(const uavca_config* c, uavo_state* s, int b, const float* action, const double* a64, int action_mode,
                               int evaluate, double* obs, double* reward, uint8_t* done) {
  const int N = c->num_agents;
  double* pos = s->pos64 + (size_t)b * N * 2;
  double* vel = s->vel + (size_t)b * N * 2;
  const double* tgt = s->tgt64 + (size_t)b * N * 2;
  const double* init = s->init64 + (size_t)b * N;
  double* prev = s->prev64 + (size_t)b * N;
  uint8_t* flags = s->flags + (size_t)b * N;
  const double tau = c->tau, amax = c->max_acceleration, vmax = c->max_speed;
  const double vm2 = n64(vmax, vmax);
  const double two_r = 2.0 * c->collider_radius, two_h = 2.0 * c->hard_collision_radius;
  const double lox = -c->x_size / 2.0, hix = c->x_size / 2.0, loy = -c->y_size / 2.0, hiy = c->y_size / 2.0;
  for (int i = 0; i < N; ++i) {
    const int parked = flags[i] & UAVCA_FLAG_PARKED;
    double prev_d, dist;
    if (parked) {
      prev_d = 0.0; dist = 0.0;
    } else {
      double ax, ay;
      if (a64) { ax = a64[2 * i]; ay = a64[2 * i + 1]; }
      else { float fx, fy; map_action(c, action_mode, action[2 * i], action[2 * i + 1], &fx, &fy); ax = (double)fx; ay = (double)fy; }
      double dvx = clipd((ax - vel[2 * i]) / tau, -amax, amax);
      double dvy = clipd((ay - vel[2 * i + 1]) / tau, -amax, amax);
      double vx = clipd(vel[2 * i] + dvx * tau, -vmax, vmax);
      double vy = clipd(vel[2 * i + 1] + dvy * tau, -vmax, vmax);
      pos[2 * i] = pos[2 * i] + vx * tau;                      /* uav_agent.py:28-29 in float64 */
      pos[2 * i + 1] = pos[2 * i + 1] + vy * tau;
      vel[2 * i] = vx; vel[2 * i + 1] = vy;
      prev_d = prev[i];
      dist = n64(tgt[2 * i] - pos[2 * i], tgt[2 * i + 1] - pos[2 * i + 1]);
    }
    const double px = pos[2 * i], py = pos[2 * i + 1];
    double dth = wrap(atan2(tgt[2 * i + 1] - py, tgt[2 * i] - px) - atan2(vel[2 * i + 1], vel[2 * i]));
    double m = vm2 / init[i];
    if (1.0 < m) m = 1.0;
    double r = 0.0 - 0.01 * m;
    r += 50.0 * ((prev_d - dist) / vm2);
    r *= (r > 0) ? (1.0 - dist / (1.5 * init[i])) : (1.0 + dist / (1.5 * init[i]));
    r -= 0.01 * fabs(dth);
    int j1, j2; double d1, d2;
    nearest2_f64(pos, N, i, &j1, &j2, &d1, &d2);
    int collision = 0;
    if (j1 >= 0 && d1 < c->d_sense) {                          /* the nearest in-range neighbour decides (:199-210) */
      if (d1 <= two_r) { r = -2.0; collision = 1; }
      if (d1 <= two_h && !parked && !(flags[i] & UAVCA_FLAG_COLLIDED)) { s->coll[b] += 1; flags[i] |= UAVCA_FLAG_COLLIDED; }
    }
    const double speed = n64(vel[2 * i], vel[2 * i + 1]);
    const int oob = !(px >= lox && px <= hix && py >= loy && py <= hiy);
    int d;
    if (dist < c->reach_distance && !collision && speed < c->reach_speed) {
      d = 1;
      if (!parked) s->reach[b] += 1;
      flags[i] |= UAVCA_FLAG_PARKED;
      double fx = vel[2 * i] / speed * 0.001, fy = vel[2 * i + 1] / speed * 0.001;
      if (fx != fx || fy != fy) { fx = 0.0; fy = 0.0; }
      vel[2 * i] = fx; vel[2 * i + 1] = fy;
      r += 10.0;
    } else if (oob) {
      d = evaluate ? 0 : 1;
    } else {
      d = 0;
    }
    prev[i] = dist;
    reward[i] = r;
    done[i] = (uint8_t)d;
  }
  for (int i = 0; i < N; ++i) obs_multi_f64(c, pos, vel, tgt, N, i, obs + 10 * i);
  s->steps[b] += 1;
  if (c->track_scores && s->score) {
    double live = 0.0;
    for (int i = 0; i < N; ++i) live += reward[i] * (1 - done[i]);
    s->score[2 * b] += reward[0];
    s->score[2 * b + 1] += live;
  }
  mirror_f32(s, b, N);
}

/* ---- persistent pthread pool: parallel-for over environments ----------------------------------------
 * Workers are created once and parked on a condition variable; a job hands out chunks of envs through an
 * atomic counter (dynamic schedule: episodes that reset cost more than ones that do not), the calling
 * thread works too.  No thread is created or joined per step, so the all-cores CPU arm of bench.py measures
 * the step itself. */

typedef void (*env_fn)(void* ctx, int b);
#define POOL_MAX 64
static struct {
  pthread_t th[POOL_MAX];
  int nthreads; /* workers created so far */
  pthread_mutex_t mu;
  pthread_cond_t cv_start, cv_done;
  unsigned long gen; /* job generation */
  int want;          /* workers that take part in the current job */
  int running;       /* workers still inside the current job */
  env_fn fn;
  void* ctx;
  int B, chunk;
  int next; /* atomic: first env of the next chunk */
} g_pool = {.mu = PTHREAD_MUTEX_INITIALIZER, .cv_start = PTHREAD_COND_INITIALIZER, .cv_done = PTHREAD_COND_INITIALIZER};

static void pool_drain(void) {
  for (;;) {
    int lo = __atomic_fetch_add(&g_pool.next, g_pool.chunk, __ATOMIC_RELAXED);
    if (lo >= g_pool.B) return;
    int hi = lo + g_pool.chunk;
    if (hi > g_pool.B) hi = g_pool.B;
    for (int b = lo; b < hi; ++b) g_pool.fn(g_pool.ctx, b);
  }
}
static void* pool_worker(void* arg) {
  const int id = (int)(intptr_t)arg;
  unsigned long seen = 0;
  pthread_mutex_lock(&g_pool.mu);
  for (;;) {
    while (g_pool.gen == seen) pthread_cond_wait(&g_pool.cv_start, &g_pool.mu);
    seen = g_pool.gen;
    if (id >= g_pool.want) continue;
    pthread_mutex_unlock(&g_pool.mu);
    pool_drain();
    pthread_mutex_lock(&g_pool.mu);
    if (--g_pool.running == 0) pthread_cond_signal(&g_pool.cv_done);
  }
  return 0;
}
static void parallel_for(env_fn fn, void* ctx, int B, int nthreads) {
  if (nthreads > POOL_MAX) nthreads = POOL_MAX;
  if (nthreads <= 1 || B < 2 * nthreads) {
    for (int b = 0; b < B; ++b) fn(ctx, b);
    return;
  }
  const int helpers = nthreads - 1; /* the caller is the last worker */
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.nthreads < helpers) {
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
    if (pthread_create(&g_pool.th[g_pool.nthreads], &at, pool_worker, (void*)(intptr_t)g_pool.nthreads) != 0) break;
    g_pool.nthreads += 1;
  }
  g_pool.fn = fn; g_pool.ctx = ctx; g_pool.B = B;
  g_pool.chunk = B / (8 * nthreads) > 0 ? B / (8 * nthreads) : 1;
  __atomic_store_n(&g_pool.next, 0, __ATOMIC_RELAXED);
  g_pool.want = helpers < g_pool.nthreads ? helpers : g_pool.nthreads;
  g_pool.running = g_pool.want;
  g_pool.gen += 1;
  pthread_cond_broadcast(&g_pool.cv_start);
  pthread_mutex_unlock(&g_pool.mu);
  pool_drain();
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.running > 0) pthread_cond_wait(&g_pool.cv_done, &g_pool.mu);
  pthread_mutex_unlock(&g_pool.mu);
}

/* ---- batched entry points --------------------------------------------------------------------------- */

typedef struct {
  const uavca_config* c; uavo_state* s; const float* action; int action_mode; int evaluate;
  double* obs; double* reward; uint8_t* done; float* distance; double* final_obs;
  const uavo_state* pool; int pool_envs; uint8_t* reset_mask;
  const double* action64; /* float64 cartesian actions instead of `action` (uavo_step_f64act) */
} step_ctx;
static int want_reset(const uavca_config* c, const uint8_t* done, int N, int steps);
static void step_multi_body(void* p, int b) {
  step_ctx* x = (step_ctx*)p;
  const int N = x->c->num_agents;
  const int f64 = x->c->circular && x->s->pos64;
  const float* af = x->action ? x->action + (size_t)b * N * 2 : 0;
  const double* ad = x->action64 ? x->action64 + (size_t)b * N * 2 : 0;
  if (f64)
    step_env_multi_f64(x->c, x->s, b, af, ad, x->action_mode, x->evaluate, x->obs + (size_t)b * N * 10,
                       x->reward + (size_t)b * N, x->done + (size_t)b * N);
  else
    step_env_multi(x->c, x->s, b, af, ad, x->action_mode, x->evaluate, x->obs + (size_t)b * N * 10,
                   x->reward + (size_t)b * N, x->done + (size_t)b * N);
  if (x->final_obs) memcpy(x->final_obs + (size_t)b * N * 10, x->obs + (size_t)b * N * 10, sizeof(double) * N * 10);
  /* auto-reset in place (envs are independent; the shared totals are folded atomically) */
  const int rs = want_reset(x->c, x->done + (size_t)b * N, N, x->s->steps[b]);
  if (x->reset_mask) x->reset_mask[b] = (uint8_t)rs;
  if (rs) {
    uavo_state* s = x->s;
    if (f64) {
      reset_env_circular(x->c, s, b);
      for (int i = 0; i < N; ++i)
        obs_multi_f64(x->c, s->pos64 + (size_t)b * N * 2, s->vel + (size_t)b * N * 2, s->tgt64 + (size_t)b * N * 2, N, i,
                      x->obs + ((size_t)b * N + i) * 10);
      return;
    }
    reset_env_multi(x->c, s, x->pool, x->pool_envs, b);
    for (int i = 0; i < N; ++i)
      obs_multi(x->c, s->pos + (size_t)b * N * 2, s->vel + (size_t)b * N * 2, s->tgt + (size_t)b * N * 2, N, i,
                x->obs + ((size_t)b * N + i) * 10);
  }
}
static void step_single_body(void* p, int b) {
  step_ctx* x = (step_ctx*)p;
  step_env_single(x->c, x->s, b, x->action ? x->action + (size_t)b * 2 : 0, x->action64 ? x->action64 + (size_t)b * 2 : 0,
                  x->action_mode, x->obs + (size_t)b * 4, x->reward + b, x->done + b, x->distance ? x->distance + b : 0);
  if (x->final_obs) memcpy(x->final_obs + (size_t)b * 4, x->obs + (size_t)b * 4, sizeof(double) * 4);
  const int rs = want_reset(x->c, x->done + b, 1, x->s->steps[b]);
  if (x->reset_mask) x->reset_mask[b] = (uint8_t)rs;
  if (rs) {
    uavo_state* s = x->s;
    reset_env_single(x->c, s, x->pool, x->pool_envs, b);
    obs_single(x->c, s->pos + (size_t)b * 2, s->vel + (size_t)b * 2, s->tgt + (size_t)b * 2, 1, x->obs + (size_t)b * 4);
  }
}

static int want_reset(const uavca_config* c, const uint8_t* done, int N, int steps) {
  int any = 0, all = 1;
  for (int i = 0; i < N; ++i) { any |= done[i]; all &= done[i]; }
  if ((c->reset_mode & UAVCA_RESET_ON_DONE0) && done[0]) return 1;
  if ((c->reset_mode & UAVCA_RESET_ON_ALL_DONE) && all) return 1;
  if ((c->reset_mode & UAVCA_RESET_ON_ANY_DONE) && any) return 1;
  if (c->max_episode_steps > 0 && steps >= c->max_episode_steps) return 1;
  return 0;
}

int uavo_reset(const uavca_config* c, uavo_state* s, const uavo_state* pool, int pool_envs, const uint8_t* mask,
               double* obs) {
  const int N = c->num_agents;
  for (int b = 0; b < c->num_envs; ++b) { /* serial: the totals in stats are shared */
    if (mask && !mask[b]) continue;
    if (c->kind == UAVCA_KIND_SINGLE) {
      reset_env_single(c, s, pool, pool_envs, b);
      if (obs) obs_single(c, s->pos + (size_t)b * 2, s->vel + (size_t)b * 2, s->tgt + (size_t)b * 2, 1, obs + (size_t)b * 4);
    } else if (c->circular && s->pos64) {
      reset_env_circular(c, s, b);
      if (obs)
        for (int i = 0; i < N; ++i)
          obs_multi_f64(c, s->pos64 + (size_t)b * N * 2, s->vel + (size_t)b * N * 2, s->tgt64 + (size_t)b * N * 2, N, i,
                        obs + ((size_t)b * N + i) * 10);
    } else {
      reset_env_multi(c, s, pool, pool_envs, b);
      if (obs)
        for (int i = 0; i < N; ++i)
          obs_multi(c, s->pos + (size_t)b * N * 2, s->vel + (size_t)b * N * 2, s->tgt + (size_t)b * N * 2, N, i,
                    obs + ((size_t)b * N + i) * 10);
    }
  }
  return 0;
}

int uavo_observe(const uavca_config* c, const uavo_state* s, double* obs) {
  const int N = c->num_agents;
  for (int b = 0; b < c->num_envs; ++b) {
    if (c->kind == UAVCA_KIND_SINGLE)
      obs_single(c, s->pos + (size_t)b * 2, s->vel + (size_t)b * 2, s->tgt + (size_t)b * 2, s->steps[b] == 0, obs + (size_t)b * 4);
    else if (c->circular && s->pos64)
      for (int i = 0; i < N; ++i)
        obs_multi_f64(c, s->pos64 + (size_t)b * N * 2, s->vel + (size_t)b * N * 2, s->tgt64 + (size_t)b * N * 2, N, i,
                      obs + ((size_t)b * N + i) * 10);
    else
      for (int i = 0; i < N; ++i)
        obs_multi(c, s->pos + (size_t)b * N * 2, s->vel + (size_t)b * N * 2, s->tgt + (size_t)b * N * 2, N, i,
                  obs + ((size_t)b * N + i) * 10);
  }
  return 0;
}

int uavo_step_multi(const uavca_config* c, uavo_state* s, const uavo_state* pool, int pool_envs, const float* action,
                    int action_mode, int evaluate, double* obs, double* reward, uint8_t* done, double* final_obs,
                    uint8_t* reset_mask, int nthreads) {
  const int N = c->num_agents, B = c->num_envs;
  step_ctx ctx = {c, s, action, action_mode, evaluate, obs, reward, done, 0, final_obs, pool, pool_envs, reset_mask, 0};
  (void)N; (void)B;
  parallel_for(step_multi_body, &ctx, c->num_envs, nthreads);
  return 0;
}

int uavo_step_single(const uavca_config* c, uavo_state* s, const uavo_state* pool, int pool_envs, const float* action,
                     int action_mode, double* obs, double* reward, uint8_t* done, float* distance, double* final_obs,
                     uint8_t* reset_mask, int nthreads) {
  const int B = c->num_envs;
  step_ctx ctx = {c, s, action, action_mode, 0, obs, reward, done, distance, final_obs, pool, pool_envs, reset_mask, 0};
  parallel_for(step_single_body, &ctx, B, nthreads);
  return 0;
}

/* Either world stepped with FLOAT64 cartesian actions (double [B][N][2]): what NumPy hands the reference's step() in its own
 * training loops (test_sac_multi.py:77-80), consumed in float64 by uav_agent.py:26 / uav_world_2d.py:142. */
int uavo_step_f64act(const uavca_config* c, uavo_state* s, const uavo_state* pool, int pool_envs, const double* action64,
                     int evaluate, double* obs, double* reward, uint8_t* done, float* distance, double* final_obs,
                     uint8_t* reset_mask, int nthreads) {
  step_ctx ctx = {c, s, 0, UAVCA_ACTION_CARTESIAN, evaluate, obs, reward, done, distance, final_obs, pool, pool_envs, reset_mask,
                  action64};
  parallel_for(c->kind == UAVCA_KIND_SINGLE ? step_single_body : step_multi_body, &ctx, c->num_envs, nthreads);
  return 0;
}

int uavo_map_action(const uavca_config* c, const float* in, int action_mode, float* out) {
  const size_t M = (size_t)c->num_envs * c->num_agents;
  for (size_t m = 0; m < M; ++m) map_action(c, action_mode, in[2 * m], in[2 * m + 1], out + 2 * m, out + 2 * m + 1);
  return 0;
}

/* The random-action stream of the driver loops (run.py:10-16, run_multi.py:10-16: env.action_space.sample()) as a
 * counter-based draw: uniform in [-1, 1)^2 per (global env, UAV, global step t); one Philox block serves steps 2q, 2q+1.
 * 24-bit uniforms, so every value is exact in float32. */
int uavo_sample_actions(const uavca_config* c, uint64_t seed, uint64_t t, float* out) {
  const int N = c->num_agents;
  for (int b = 0; b < c->num_envs; ++b)
    for (int i = 0; i < N; ++i) {
      uint32_t r[4];
      philox4x32_10((uint32_t)(c->env_index_base + b), (uint32_t)(t >> 1), ((uint32_t)STREAM_ACT << 16) | (uint32_t)i,
                    (uint32_t)(t >> 33), (uint32_t)seed, (uint32_t)(seed >> 32), r);
      const uint32_t wa = (t & 1) ? r[2] : r[0], wb = (t & 1) ? r[3] : r[1];
      out[((size_t)b * N + i) * 2] = (float)((double)(wa >> 8) * (2.0 / 16777216.0) - 1.0);
      out[((size_t)b * N + i) * 2 + 1] = (float)((double)(wb >> 8) * (2.0 / 16777216.0) - 1.0);
    }
  return 0;
}

int uavo_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (n > 64 ? 64 : (int)n);
}
