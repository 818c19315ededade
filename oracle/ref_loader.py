"""Import the UNMODIFIED reference environment classes in this container.  TEST INFRASTRUCTURE ONLY.

The reference (``/root/reference``) imports ``gym``, ``pygame`` and ``turtle``, none of which is installed
here; its environment core uses nothing from them beyond ``gym.Env`` as a base class and ``spaces.Box`` as
a record.  This module registers minimal stand-ins in ``sys.modules`` and imports

    gym_uav_collision_avoidance.envs.UAVWorld2D        (uav_world_2d.py:11)
    gym_uav_collision_avoidance.envs.MultiUAVWorld2D   (multi_uav_world_2d.py:10)

from ``/root/reference`` untouched — or, on the GPU box where that checkout does not exist, from the unmodified
copies of the five env-core files that oracle/build_ref.py leaves under ``oracle/_ref/`` (git-ignored).  In the
build container it validates oracle/uav_oracle.c and generates tests/golden/; on the GPU box it only serves
bench.py's literal-reference CPU legs.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("UAVCA_REFERENCE_ROOT", "/root/reference")
_TRAVELLING_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # made by oracle/build_ref.py


def _has_envs(root: str) -> bool:
    return os.path.isdir(os.path.join(root, "gym_uav_collision_avoidance", "envs"))


if not _has_envs(REFERENCE_ROOT) and _has_envs(_TRAVELLING_COPY):
    REFERENCE_ROOT = _TRAVELLING_COPY  # the GPU box: unmodified copies of the five env-core files


def reference_available() -> bool:
    return _has_envs(REFERENCE_ROOT)


def reference_is_checkout() -> bool:
    """True when the reference is the original checkout (the build container), not the travelling copy."""
    return reference_available() and REFERENCE_ROOT != _TRAVELLING_COPY


class _Box:
    """The subset of gym.spaces.Box the reference and its callers touch."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def sample(self):
        return np.random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)


def _install_stubs() -> None:
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # gym.Env
            metadata: dict = {}

            def close(self):
                pass

        gym.Env = Env
        spaces = types.ModuleType("gym.spaces")
        spaces.Box = _Box
        gym.spaces = spaces
        envs = types.ModuleType("gym.envs")
        registration = types.ModuleType("gym.envs.registration")
        registration.registry = {}

        def register(id, entry_point=None, **kwargs):  # noqa: A002
            registration.registry[id] = entry_point

        registration.register = register
        envs.registration = registration
        gym.envs = envs
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
        sys.modules["gym.envs"] = envs
        sys.modules["gym.envs.registration"] = registration
    if "pygame" not in sys.modules:
        sys.modules["pygame"] = types.ModuleType("pygame")
    if "turtle" not in sys.modules:
        turtle = types.ModuleType("turtle")
        turtle.position = lambda *a, **k: None
        sys.modules["turtle"] = turtle
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            cv2 = types.ModuleType("cv2")
            cv2.normalize = lambda *a, **k: None
            cv2.resizeWindow = lambda *a, **k: None
            sys.modules["cv2"] = cv2


def load_reference():
    """Return (UAVWorld2D, MultiUAVWorld2D) imported from the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from gym_uav_collision_avoidance.envs import MultiUAVWorld2D, UAVWorld2D  # type: ignore

    return UAVWorld2D, MultiUAVWorld2D


# ---- state injection / extraction ---------------------------------------------------------


def inject_multi(env, pos, vel, tgt, init, prev, flags, steps=0, reach=0, coll=0):
    """Write one env's SoA state into a literal MultiUAVWorld2D (arrays of shape [N,2] / [N])."""
    for i, a in enumerate(env.agent_list):
        a.location = np.array(pos[i], dtype=np.float32)  # fresh: mutated in place (uav_agent.py:29)
        a.velocity = np.array(vel[i], dtype=np.float64)
        a.velocity_prev = np.array(vel[i], dtype=np.float64)
        a.target_location = np.array(tgt[i], dtype=np.float32)
        a.init_distance = np.float32(init[i])
        a.prev_distance = np.float32(prev[i])
        a.done = bool(flags[i] & 1)
        a.collided = bool(flags[i] & 2)
    env.steps = int(steps)
    env.target_reach_count = int(reach)
    env.collision_count = int(coll)


def extract_multi(env):
    n = len(env.agent_list)
    pos = np.stack([np.asarray(a.location, dtype=np.float32) for a in env.agent_list])
    vel = np.stack([np.asarray(a.velocity, dtype=np.float64) for a in env.agent_list])
    tgt = np.stack([np.asarray(a.target_location, dtype=np.float32) for a in env.agent_list])
    init = np.array([a.init_distance for a in env.agent_list], dtype=np.float32)
    prev = np.array([a.prev_distance for a in env.agent_list], dtype=np.float32)
    flags = np.array([(1 if a.done else 0) | (2 if a.collided else 0) for a in env.agent_list], dtype=np.uint8)
    assert pos.shape == (n, 2)
    return pos, vel, tgt, init, prev, flags


def inject_single(env, pos, vel, tgt, init, prev, steps=0, vel_f32=False):
    env._agent_location = np.array(pos, dtype=np.float32)
    vdt = np.float32 if vel_f32 else np.float64
    env._agent_speed = np.array(vel, dtype=vdt)
    env._agent_speed_prev = env._agent_speed
    env._target_location = np.array(tgt, dtype=np.float32)
    env._init_target_distance = np.float32(init)
    env._prev_distance = np.float32(prev)
    env.steps = int(steps)


def extract_single(env):
    return (
        np.asarray(env._agent_location, dtype=np.float32).copy(),
        np.asarray(env._agent_speed, dtype=np.float64).copy(),
        np.asarray(env._target_location, dtype=np.float32).copy(),
        np.float32(env._init_target_distance),
        np.float32(env._prev_distance),
    )
