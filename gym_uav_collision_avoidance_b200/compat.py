"""B=1 drop-ins for the reference classes: same constructors, return types and attributes, GPU step underneath.

`MultiUAVWorld2D` / `UAVWorld2D` keep the reference's gym 0.24 contract (multi_uav_world_2d.py:116,177;
uav_world_2d.py:119,137): `reset()` / `step()` take and return Python lists / NumPy arrays (float64 observations,
4-tuple `obs, reward, done, info`), `agent_list[i]` exposes `location`, `target_location`, `velocity`, `done`, ...
(readable and assignable, as test_sac_multi_plot_trajectory.py:46-47 and test_ddpg_multi.py:122-126 use them), and
the counters `steps`, `target_reach_count`, `collision_count` live on the env object.  They exist so that the
reference's loops run unchanged against the CUDA path (parity, migration); throughput comes from the batched
classes in `batched.py`.  Differences, all forced by the design: episodes are drawn from the counter-based Philox
stream (`seed=` kwarg) instead of the global `np.random`; observations/rewards are float32 values widened to
float64; in-place edits of a returned array (`agent.location[0] = x`) do not reach the device — assign the
attribute instead; `render()` opens no window but records the frame into `env.trajectory` (`export_trajectory()`).
Actions keep their dtype: python floats / float64 arrays (what the SAC/TD3/DDPG loops build, test_sac_multi.py:77-80) go
through `uavca_step_f64` and are consumed in float64 like the reference's `(action - v) / tau` (uav_agent.py:26), float32
arrays (`action_space.sample()`) through the float32 entry points.
"""
from __future__ import annotations

import numpy as np
import torch

from .batched import BatchedMultiUAVWorld2D, BatchedUAVWorld2D

GYM_IDS = {  # gym_uav_collision_avoidance/__init__.py:3-10
    "gym_uav_collision_avoidance/UAVWorld2D-v0": "gym_uav_collision_avoidance_b200.compat:UAVWorld2D",
    "gym_uav_collision_avoidance/MultiUAVWorld2D-v0": "gym_uav_collision_avoidance_b200.compat:MultiUAVWorld2D",
}


def register_gym_ids() -> bool:
    """Register the reference's ids with gym / gymnasium when one of them is importable."""
    for mod in ("gym", "gymnasium"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"]).register
        except Exception:
            continue
        for gid, entry in GYM_IDS.items():
            try:
                reg(id=gid, entry_point=entry)
            except Exception:
                pass
        return True
    return False


def _raw_stream(device) -> int:
    """torch's current stream on `device` as a cudaStream_t (the raw getter costs a fifth of building a Stream object)."""
    try:
        return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream(device).cuda_stream


class _HostIO:
    """Pinned host tensors the B=1 step reads its action from and writes obs / reward / done into directly (mapped host
    memory: the kernel's loads and stores go through PCIe, one launch + one stream synchronisation per step, no copies)."""

    def __init__(self, b):
        N, D = b.num_agents, b.obs_dim
        pin = lambda *shape, dtype=torch.float32: torch.zeros(shape, dtype=dtype).pin_memory()  # noqa: E731
        self.action, self.obs, self.reward = pin(1, N, 2), pin(1, N, D), pin(1, N)
        self.done, self.mask, self.distance = pin(1, N, dtype=torch.uint8), pin(1, dtype=torch.uint8), pin(1)
        self.action64 = pin(1, N, 2, dtype=torch.float64)
        self.np = {k: getattr(self, k).numpy() for k in ("action", "action64", "obs", "reward", "done", "distance")}
        self.ptr = {k: getattr(self, k).data_ptr() for k in ("action", "action64", "obs", "reward", "done", "mask", "distance")}
        self._step_sync = b._lib.uavca_step_sync
        self._single = b.kind == 1  # UAVCA_KIND_SINGLE: `distance` is an output

    def step(self, b, n_action, evaluate=False):
        """One `env.step`: the action into the mapped buffer of its dtype, ONE C call that launches the step and waits for
        the stream (`uavca_step_sync`), results readable in the mapped output buffers."""
        n_action = np.asarray(n_action)
        p = self.ptr
        f64 = n_action.dtype != np.float32
        if f64:
            # what the reference's loops hand env.step(): python floats / float64 arrays (test_sac_multi.py:77-80), consumed in
            # float64 by UAVAgent.step (uav_agent.py:26) — kept unrounded
            self.np["action64"][...] = n_action.reshape(1, b.num_agents, 2)
        else:
            self.np["action"][...] = n_action.reshape(1, b.num_agents, 2)
        rc = self._step_sync(b._h, b.state.blob.data_ptr(), p["action64"] if f64 else p["action"], int(f64), 0, int(bool(evaluate)), p["obs"],
                             p["reward"], p["done"], p["distance"] if self._single else None, None, p["mask"], _raw_stream(b.device))
        if rc:
            from . import _capi

            _capi.check(rc, "uavca_step_sync")
        return self.np


class _AgentView:
    """`UAVAgent` (uav_agent.py:7-20) as a window onto UAV `i` of the device state."""

    def __init__(self, env: "MultiUAVWorld2D", i: int, color):
        self._env, self._i, self.color = env, i, color
        self.max_speed = env.max_speed.copy()
        self.max_acceleration = env.max_acceleratoin.copy()
        self.tau = env.tau

    _F64 = {"pos": "pos64", "tgt": "tgt64", "init": "init64", "prev": "prev64"}

    def _field(self, field):
        st = self._env._live.state
        if st.float64_world:  # an episode started by reset(circular=True): the reference holds float64 arrays there
            field = self._F64.get(field, field)
        return getattr(st, field)

    def _get(self, field):
        return self._field(field)[0, self._i].cpu().numpy()

    def _set(self, field, value):
        t = self._field(field)
        t[0, self._i] = torch.as_tensor(np.asarray(value, dtype=np.float64)).to(t.dtype)
        st = self._env._live.state
        if st.float64_world and field in self._F64:  # keep the float32 mirror in step
            m = getattr(st, field)
            m[0, self._i] = t[0, self._i].to(m.dtype)

    location = property(lambda s: s._get("pos"), lambda s, v: s._set("pos", v))
    target_location = property(lambda s: s._get("tgt"), lambda s, v: s._set("tgt", v))
    velocity = property(lambda s: s._get("vel"), lambda s, v: s._set("vel", v))
    velocity_prev = property(lambda s: s._get("vel"), lambda s, v: s._set("vel", v))  # both names hold v after a step
    init_distance = property(lambda s: np.float32(s._get("init")), lambda s, v: s._set("init", v))
    prev_distance = property(lambda s: np.float32(s._get("prev")), lambda s, v: s._set("prev", v))

    def _flag(self, bit):
        return bool(int(self._env._live.state.flags[0, self._i].item()) & bit)

    def _set_flag(self, bit, on):
        f = self._env._live.state.flags
        cur = int(f[0, self._i].item())
        f[0, self._i] = (cur | bit) if on else (cur & ~bit)

    done = property(lambda s: s._flag(1), lambda s, v: s._set_flag(1, bool(v)))
    collided = property(lambda s: s._flag(2), lambda s, v: s._set_flag(2, bool(v)))


class MultiUAVWorld2D:
    metadata = {"render_fps": 1000}

    def __init__(self, x_size=50.0, y_size=50.0, max_speed=10.0, max_acceleration=5.0, num_agents=4, collider_radius=1.0,
                 d_sense=15, *, seed=0, device=None):
        import colorsys

        kw = dict(x_size=x_size, y_size=y_size, max_speed=max_speed, max_acceleration=max_acceleration,
                  num_agents=num_agents, collider_radius=collider_radius, d_sense=d_sense, device=device, seed=seed)
        self._b = BatchedMultiUAVWorld2D(1, **kw)
        self._kw = kw
        self._io = _HostIO(self._b)
        self._circ = None  # float64-world twin (its own state), created on the first reset(circular=True)
        self._live = self._b  # the env the current episode runs in
        for name in ("x_size", "y_size", "map_diagonal_size", "min_location", "max_location", "max_speed", "min_speed",
                     "max_acceleratoin", "min_acceleratoin", "tau", "collider_radius", "d_sense", "observation_space",
                     "action_space"):
            setattr(self, name, getattr(self._b, name))
        self.num_agents = num_agents
        self.map_dimension = np.array([x_size, y_size])
        self.agent_list = []
        for i in range(num_agents):  # multi_uav_world_2d.py:36-41
            r, g, b = colorsys.hsv_to_rgb(i / num_agents, 1.0, 1.0)
            self.agent_list.append(_AgentView(self, i, (int(255 * r), int(255 * g), int(255 * b))))
        self.window = self.clock = None
        self.trajectory = []  # frames recorded by render()

    # counters live in the device state (multi_uav_world_2d.py:166-168)
    def _counter(name):
        def get(self):
            return int(getattr(self._live.state, name)[0].item())

        def set_(self, v):
            getattr(self._live.state, name)[0] = int(v)

        return property(get, set_)

    steps = _counter("steps")
    target_reach_count = _counter("reach")
    collision_count = _counter("coll")
    del _counter

    def _obs_list(self, obs: torch.Tensor):
        o = obs[0].cpu().numpy().astype(np.float64)
        return [o[i] for i in range(self.num_agents)]

    def _get_obs(self, agent):
        return self._obs_list(self._live.observe())[agent._i]

    def reset(self, return_info=False, circular=False):
        """reset(circular=True) starts an episode in the FLOAT64 world (multi_uav_world_2d.py:157-163 assigns float64
        arrays to the locations, and they stay float64 until the next plain reset()): it runs on a twin env whose state
        keeps float64 positions (`circular=True` of the batched class)."""
        if circular:
            if self._circ is None:
                self._circ = BatchedMultiUAVWorld2D(1, circular=True, **self._kw)
            self._live = self._circ
        else:
            self._live = self._b
        obs = self._obs_list(self._live.reset())
        info = {"distance": 0}
        return (obs, info) if return_info else obs

    def step(self, n_action, evaluate=False):
        out = self._io.step(self._live, n_action, evaluate)
        o = out["obs"][0].astype(np.float64)
        return ([o[i] for i in range(self.num_agents)], [float(x) for x in out["reward"][0]],
                [bool(x) for x in out["done"][0]], {"distance": 0})

    def render(self, mode="human"):
        """No window: the reference's pygame drawing (multi_uav_world_2d.py:243-331) is replaced by a trajectory tap.
        Every call appends the current locations / targets / done latches to `self.trajectory` (a list of dicts of
        NumPy arrays), which is what the reference's plotting script collects by hand
        (test_sac_multi_plot_trajectory.py:46-68); `export_trajectory()` stacks it."""
        st = self._live.state
        self.trajectory.append(dict(pos=st.pos[0].cpu().numpy(), target=st.tgt[0].cpu().numpy(),
                                    done=(st.flags[0].cpu().numpy() & 1).astype(bool), step=self.steps))
        return None

    def export_trajectory(self) -> dict:
        """pos [T, N, 2], target [T, N, 2], done [T, N], step [T] of the frames recorded by render()."""
        if not self.trajectory:
            return dict(pos=np.zeros((0, self.num_agents, 2), np.float32), target=np.zeros((0, self.num_agents, 2), np.float32),
                        done=np.zeros((0, self.num_agents), bool), step=np.zeros(0, np.int64))
        return {k: np.stack([f[k] for f in self.trajectory]) for k in ("pos", "target", "done", "step")}

    def close(self):
        self._b.close()
        if self._circ is not None:
            self._circ.close()


class UAVWorld2D:
    metadata = {"render_fps": 1000}

    def __init__(self, x_size=100.0, y_size=100.0, agent_num=4, max_speed=12.0, max_acceleration=5.0, *, seed=0,
                 device=None, float32_actions=False):
        # float32_actions: the caller feeds float32 arrays (`action_space.sample()`, run.py:11); the reference then forms
        # the first quotient after a reset in float32 (uav_world_2d.py:122,142).  The SAC/TD3 loops feed float64.
        self._b = BatchedUAVWorld2D(1, x_size=x_size, y_size=y_size, max_speed=max_speed, max_acceleration=max_acceleration,
                                    device=device, seed=seed, float32_first_step=float32_actions)
        for name in ("x_size", "y_size", "map_diagonal_size", "min_location", "max_location", "max_speed", "min_speed",
                     "max_acceleratoin", "min_acceleratoin", "tau", "observation_space", "action_space"):
            setattr(self, name, getattr(self._b, name))
        self.map_dimension = np.array([x_size, y_size])
        self.window = self.clock = None
        self._io = _HostIO(self._b)

    def _field(name, scalar=False):
        def get(self):
            v = getattr(self._b.state, name)[0, 0].cpu().numpy()
            return np.float32(v) if scalar else v

        def set_(self, v):
            t = getattr(self._b.state, name)
            t[0, 0] = torch.as_tensor(np.asarray(v, dtype=np.float64)).to(t.dtype)

        return property(get, set_)

    _agent_location = _field("pos")
    _target_location = _field("tgt")
    _agent_speed = _field("vel")
    _agent_speed_prev = _field("vel")
    _init_target_distance = _field("init", True)
    _prev_distance = _field("prev", True)
    del _field

    @property
    def steps(self):
        return int(self._b.state.steps[0].item())

    @steps.setter
    def steps(self, v):
        self._b.state.steps[0] = int(v)

    def _get_obs(self):
        return self._b.observe()[0, 0].cpu().numpy().astype(np.float64)

    def _get_info(self):  # uav_world_2d.py:114-117
        return {"distance": np.float32(np.linalg.norm(self._target_location - self._agent_location))}

    def reset(self, return_info=False, options=None):
        obs = self._b.reset()[0, 0].cpu().numpy().astype(np.float64)
        return (obs, self._get_info()) if return_info else obs

    def step(self, action):
        out = self._io.step(self._b, action)
        return (out["obs"][0, 0].astype(np.float64), np.float32(out["reward"][0, 0]), bool(out["done"][0, 0]),
                {"distance": np.float32(out["distance"][0])})

    def render(self, mode="human"):
        return None

    def close(self):
        self._b.close()
