"""ctypes front-end of oracle/uav_oracle.c.  TEST INFRASTRUCTURE ONLY (see the header of uav_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
State lives in NumPy arrays with the same structure-of-arrays fields as the device state blob.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libuav_oracle.so")

KIND_MULTI, KIND_SINGLE = 0, 1
ACTION_CARTESIAN, ACTION_POLAR, ACTION_SCALED = 0, 1, 2
RESET_ON_DONE0, RESET_ON_ALL_DONE, RESET_ON_ANY_DONE = 1, 2, 4
SOURCE_PHILOX, SOURCE_POOL = 0, 1
FLAG_PARKED, FLAG_COLLIDED = 1, 2


class Config(C.Structure):
    """Mirror of `uavca_config` (include/uavca.h)."""

    _fields_ = [
        ("kind", C.c_int32),
        ("num_envs", C.c_int32),
        ("num_agents", C.c_int32),
        ("reset_mode", C.c_int32),
        ("max_episode_steps", C.c_int32),
        ("reset_source", C.c_int32),
        ("circular", C.c_int32),
        ("single_f32_first_step", C.c_int32),
        ("env_index_base", C.c_int64),
        ("seed", C.c_uint64),
        ("x_size", C.c_double),
        ("y_size", C.c_double),
        ("max_speed", C.c_double),
        ("max_acceleration", C.c_double),
        ("tau", C.c_double),
        ("collider_radius", C.c_double),
        ("hard_collision_radius", C.c_double),
        ("d_sense", C.c_double),
        ("reach_distance", C.c_double),
        ("reach_speed", C.c_double),
        ("polar_scale", C.c_double),
        ("track_scores", C.c_int32),
        ("reserved0", C.c_int32),
    ]


def multi_config(num_envs, num_agents, **kw) -> Config:
    """Reference defaults of MultiUAVWorld2D (multi_uav_world_2d.py:13,26,8)."""
    c = Config(kind=KIND_MULTI, num_envs=num_envs, num_agents=num_agents, x_size=50.0, y_size=50.0, max_speed=10.0,
               max_acceleration=5.0, tau=0.02, collider_radius=1.0, hard_collision_radius=0.5, d_sense=15.0,
               reach_distance=0.5, reach_speed=0.2)
    for k, v in kw.items():
        setattr(c, k, v)
    if "polar_scale" not in kw:
        # np.linalg.norm(env.action_space.high): float32 norm of [max_speed, max_speed] (test_sac_multi.py:77)
        c.polar_scale = float(np.linalg.norm(np.full(2, c.max_speed, dtype=np.float32)))
    return c


def single_config(num_envs, **kw) -> Config:
    """Reference defaults of UAVWorld2D (uav_world_2d.py:14,26)."""
    c = Config(kind=KIND_SINGLE, num_envs=num_envs, num_agents=1, x_size=100.0, y_size=100.0, max_speed=12.0,
               max_acceleration=5.0, tau=0.02, collider_radius=1.0, hard_collision_radius=0.5, d_sense=15.0,
               reach_distance=0.5, reach_speed=0.2)
    for k, v in kw.items():
        setattr(c, k, v)
    if "polar_scale" not in kw:
        c.polar_scale = float(np.float32(c.max_speed))  # env.action_space.high[0] (test_sac.py:77)
    return c


class _CState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("pos", "vel", "tgt", "init", "prev", "flags", "steps", "reach", "coll", "episode", "score", "pos64", "tgt64", "init64", "prev64", "stats")]


def build(force: bool = False) -> str:
    """Compile oracle/uav_oracle.c with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "uav_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "uavca.h")
    stale = (not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libuav_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        for name in ("uavo_reset", "uavo_observe", "uavo_step_multi", "uavo_step_single", "uavo_map_action",
                     "uavo_max_threads", "uavo_sample_actions", "uavo_step_f64act"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().uavo_max_threads())


class State:
    """Host structure-of-arrays state for B envs x N UAVs."""

    FIELDS = ("pos", "vel", "tgt", "init", "prev", "flags", "steps", "reach", "coll", "episode", "score", "pos64", "tgt64",
              "init64", "prev64", "stats")

    def __init__(self, num_envs: int, num_agents: int):
        B, N = num_envs, num_agents
        self.B, self.N = B, N
        self.pos = np.zeros((B, N, 2), np.float32)
        self.vel = np.zeros((B, N, 2), np.float64)
        self.tgt = np.zeros((B, N, 2), np.float32)
        self.init = np.ones((B, N), np.float32)
        self.prev = np.ones((B, N), np.float32)
        self.flags = np.zeros((B, N), np.uint8)
        self.steps = np.zeros(B, np.int32)
        self.reach = np.zeros(B, np.int32)
        self.coll = np.zeros(B, np.int32)
        self.episode = np.zeros(B, np.uint32)
        self.stats = np.zeros(8, np.uint64)  # [4], [5] hold float64 score sums (view with .view(np.float64))
        self.score = np.zeros((B, 2), np.float64)
        # float64 world (config.circular: the reference keeps float64 locations after reset(circular=True))
        self.pos64 = np.zeros((B, N, 2), np.float64)
        self.tgt64 = np.zeros((B, N, 2), np.float64)
        self.init64 = np.ones((B, N), np.float64)
        self.prev64 = np.ones((B, N), np.float64)

    def copy(self) -> "State":
        s = State(self.B, self.N)
        for f in self.FIELDS:
            getattr(s, f)[...] = getattr(self, f)
        return s

    def take(self, idx) -> "State":
        idx = np.asarray(idx)
        s = State(len(idx), self.N)
        for f in self.FIELDS[:-1]:
            getattr(s, f)[...] = getattr(self, f)[idx]
        return s

    def _c(self) -> _CState:
        for f in self.FIELDS:
            a = getattr(self, f)
            assert a.flags["C_CONTIGUOUS"], f
        return _CState(*[getattr(self, f).ctypes.data for f in self.FIELDS])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Batched CPU oracle: same call surface as the CUDA env (reset / observe / step)."""

    def __init__(self, cfg: Config, nthreads: int = 1):
        self.cfg = cfg
        self.B, self.N = cfg.num_envs, cfg.num_agents
        self.obs_dim = 4 if cfg.kind == KIND_SINGLE else 10
        self.state = State(self.B, self.N)
        self.pool: State | None = None
        self.nthreads = nthreads

    def set_pool(self, pool: State | None):
        self.pool = pool

    def _pool_args(self):
        if self.pool is None:
            return None, 0
        self._pool_c = self.pool._c()
        return C.byref(self._pool_c), self.pool.B

    def reset(self, mask=None):
        obs = np.zeros((self.B, self.N, self.obs_dim), np.float64)
        st = self.state._c()
        pp, pn = self._pool_args()
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().uavo_reset(C.byref(self.cfg), C.byref(st), pp, C.c_int(pn), _ptr(m), _ptr(obs))
        return obs

    def observe(self):
        obs = np.zeros((self.B, self.N, self.obs_dim), np.float64)
        st = self.state._c()
        lib().uavo_observe(C.byref(self.cfg), C.byref(st), _ptr(obs))
        return obs

    def map_action(self, action, action_mode):
        a = np.ascontiguousarray(action, dtype=np.float32).reshape(self.B, self.N, 2)
        out = np.empty_like(a)
        lib().uavo_map_action(C.byref(self.cfg), _ptr(a), C.c_int(action_mode), _ptr(out))
        return out

    def sample_actions(self, step: int, seed: int = 0):
        """Policy-space actions in [-1, 1)^2 of global step `step` (the rollout's Philox action stream)."""
        out = np.empty((self.B, self.N, 2), np.float32)
        lib().uavo_sample_actions(C.byref(self.cfg), C.c_uint64(seed), C.c_uint64(step), _ptr(out))
        return out

    def step(self, action, action_mode=ACTION_CARTESIAN, evaluate=False, want_final_obs=False):
        """Returns dict(obs f64 [B,N,D], reward f64 [B,N], done u8 [B,N], reset_mask u8 [B], ...)."""
        B, N, D = self.B, self.N, self.obs_dim
        a = np.ascontiguousarray(action, dtype=np.float32).reshape(B, N, 2)
        obs = np.zeros((B, N, D), np.float64)
        reward = np.zeros((B, N), np.float64)
        done = np.zeros((B, N), np.uint8)
        final_obs = np.zeros((B, N, D), np.float64) if want_final_obs else None
        reset_mask = np.zeros(B, np.uint8)
        st = self.state._c()
        pp, pn = self._pool_args()
        out = dict(obs=obs, reward=reward, done=done, reset_mask=reset_mask, final_obs=final_obs)
        if self.cfg.kind == KIND_SINGLE:
            dist = np.zeros(B, np.float32)
            lib().uavo_step_single(C.byref(self.cfg), C.byref(st), pp, C.c_int(pn), _ptr(a), C.c_int(action_mode),
                                   _ptr(obs), _ptr(reward), _ptr(done), _ptr(dist), _ptr(final_obs), _ptr(reset_mask),
                                   C.c_int(self.nthreads))
            out["distance"] = dist
        else:
            lib().uavo_step_multi(C.byref(self.cfg), C.byref(st), pp, C.c_int(pn), _ptr(a), C.c_int(action_mode),
                                  C.c_int(1 if evaluate else 0), _ptr(obs), _ptr(reward), _ptr(done), _ptr(final_obs),
                                  _ptr(reset_mask), C.c_int(self.nthreads))
        return out

    def step_f64(self, action64, evaluate=False, want_final_obs=False):
        """step() fed float64 cartesian actions [B,N,2] (what the reference's own loops build, test_sac_multi.py:77-80)."""
        B, N, D = self.B, self.N, self.obs_dim
        a = np.ascontiguousarray(action64, dtype=np.float64).reshape(B, N, 2)
        obs = np.zeros((B, N, D), np.float64)
        reward = np.zeros((B, N), np.float64)
        done = np.zeros((B, N), np.uint8)
        final_obs = np.zeros((B, N, D), np.float64) if want_final_obs else None
        reset_mask = np.zeros(B, np.uint8)
        dist = np.zeros(B, np.float32)
        st = self.state._c()
        pp, pn = self._pool_args()
        lib().uavo_step_f64act(C.byref(self.cfg), C.byref(st), pp, C.c_int(pn), _ptr(a), C.c_int(1 if evaluate else 0),
                               _ptr(obs), _ptr(reward), _ptr(done), _ptr(dist), _ptr(final_obs), _ptr(reset_mask),
                               C.c_int(self.nthreads))
        out = dict(obs=obs, reward=reward, done=done, reset_mask=reset_mask, final_obs=final_obs)
        if self.cfg.kind == KIND_SINGLE:
            out["distance"] = dist
        return out


# ---- synthetic initial states drawn from the reference's reset distribution (host side, NumPy) --------------


def sample_multi_states(num_envs, num_agents, rng, x_size=50.0, y_size=50.0, collider_radius=1.0, region=None) -> State:
    """Host implementation of the reset constraints (multi_uav_world_2d.py:126-155) with a NumPy generator.

    `region` (half-width, optional) confines positions/targets to a smaller square to provoke collisions."""
    st = State(num_envs, num_agents)
    hx, hy = (x_size / 2, y_size / 2) if region is None else (region, region)
    two_r = np.float32(2 * collider_radius)

    def norm32(d):
        d = d.astype(np.float32)
        return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1], dtype=np.float32)

    for b in range(num_envs):
        for i in range(num_agents):
            while True:
                p = rng.uniform([-hx, -hy], [hx, hy]).astype(np.float32)
                if i == 0 or not (norm32(st.pos[b, :i] - p) <= two_r).any():
                    st.pos[b, i] = p
                    break
        for i in range(num_agents):
            while True:
                t = rng.uniform([-hx, -hy], [hx, hy]).astype(np.float32)
                if norm32(t - st.pos[b, i]) <= two_r:
                    continue
                if i > 0 and (norm32(st.tgt[b, :i] - t) <= two_r).any():
                    continue
                st.tgt[b, i] = t
                break
    d = st.tgt - st.pos
    st.init[...] = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1], dtype=np.float32)
    st.prev[...] = st.init
    return st


def sample_single_states(num_envs, rng, x_size=100.0, y_size=100.0, max_speed=12.0) -> State:
    """Reset distribution of UAVWorld2D (uav_world_2d.py:121-130)."""
    st = State(num_envs, 1)
    hx, hy = x_size / 2, y_size / 2
    st.pos[:, 0] = rng.uniform([-hx, -hy], [hx, hy], size=(num_envs, 2)).astype(np.float32)
    st.vel[:, 0] = rng.uniform(-max_speed, max_speed, size=(num_envs, 2)).astype(np.float32).astype(np.float64)
    st.tgt[:, 0] = rng.uniform([-hx, -hy], [hx, hy], size=(num_envs, 2)).astype(np.float32)
    d = st.tgt - st.pos
    st.init[...] = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1], dtype=np.float32)
    st.prev[...] = st.init
    return st
