"""CPU: the C-ABI library builds, loads and exports every symbol include/uavca.h declares; configuration and
layout logic (no kernel launches); the product fails loudly without a GPU and never touches the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from gym_uav_collision_avoidance_b200 import _capi, build

    build.build()
    return _capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "uavca.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(uavca_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(capi):
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/uavca.h but not exported by libuavca.so"
        assert n in capi.SYMBOLS, f"{n} has no ctypes prototype in _capi.SYMBOLS"
    assert sorted(capi.SYMBOLS) == names
    import re

    header = open(os.path.join(ROOT, "include", "uavca.h")).read()
    assert lib.uavca_version() == int(re.search(r"#define\s+UAVCA_VERSION\s+(\d+)", header).group(1))
    # the driver's build check compares the same two numbers: no version literal may hide in it
    assert "uavca_version() ==" not in open(os.path.join(ROOT, "__graft_entry__.py")).read()


def test_default_configs_follow_the_reference_constructors(capi):
    m = capi.default_config(capi.KIND_MULTI)  # multi_uav_world_2d.py:13,26,8
    assert (m.x_size, m.y_size, m.max_speed, m.max_acceleration, m.num_agents) == (50.0, 50.0, 10.0, 5.0, 4)
    assert (m.collider_radius, m.d_sense, m.tau, m.hard_collision_radius) == (1.0, 15.0, 0.02, 0.5)
    assert abs(m.polar_scale - 200 ** 0.5) < 1e-6
    s = capi.default_config(capi.KIND_SINGLE)  # uav_world_2d.py:14,26
    assert (s.x_size, s.y_size, s.max_speed, s.max_acceleration, s.num_agents, s.tau) == (100.0, 100.0, 12.0, 5.0, 1, 0.02)


def test_config_struct_matches_the_oracle_mirror(capi):
    from oracle import oracle as O

    assert C.sizeof(capi.Config) == C.sizeof(O.Config)
    assert [f[0] for f in capi.Config._fields_] == [f[0] for f in O.Config._fields_]
    for (n1, t1), (n2, t2) in zip(capi.Config._fields_, O.Config._fields_):
        assert getattr(capi.Config, n1).offset == getattr(O.Config, n2).offset


def test_create_fails_loudly_without_a_gpu(capi):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = capi.default_config(capi.KIND_MULTI)
    h = C.c_void_p()
    rc = capi.load().uavca_create(C.byref(cfg), 0, C.byref(h))
    assert rc != 0 and "no CPU fallback" in capi.last_error()
    import gym_uav_collision_avoidance_b200 as G

    with pytest.raises(G.UavcaError):
        G.BatchedMultiUAVWorld2D(8, num_agents=4)


def test_bad_configs_are_rejected_before_touching_the_device(capi):
    lib = capi.load()
    for field, value in (("num_agents", 1025), ("num_agents", 0), ("num_envs", 0), ("tau", 0.0), ("kind", 7)):
        cfg = capi.default_config(capi.KIND_MULTI)
        setattr(cfg, field, value)
        h = C.c_void_p()
        assert lib.uavca_create(C.byref(cfg), 0, C.byref(h)) == -1, field
        assert capi.last_error()
    cfg = capi.default_config(capi.KIND_SINGLE)
    cfg.num_agents = 2
    h = C.c_void_p()
    assert lib.uavca_create(C.byref(cfg), 0, C.byref(h)) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gym_uav_collision_avoidance_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_misaligned_buffers_are_rejected_before_any_launch(capi):
    """The argument checks of the step entry points run before the device is touched (a fake non-null handle would be
    dereferenced, so this goes through null-handle / null-argument ordering only): null handle first."""
    lib = capi.load()
    assert lib.uavca_step_multi(None, 16, 16, 0, 0, 16, 16, 16, None, None, None) == -1
    assert "null handle" in capi.last_error()


def test_register_gym_ids_against_a_stub_gym(monkeypatch):
    """`register_gym_ids()` registers the reference's two ids (gym_uav_collision_avoidance/__init__.py:3-10) with whichever
    of gym / gymnasium is importable; neither is installed here, so a stub `gym` records the calls."""
    import sys
    import types

    from gym_uav_collision_avoidance_b200 import compat

    # drop gym / gymnasium and their submodules (oracle/ref_loader.py leaves a stub `gym.envs.registration` behind when
    # an oracle test ran first in this process)
    for name in [m for m in sys.modules if m.split(".")[0] in ("gym", "gymnasium")]:
        monkeypatch.delitem(sys.modules, name, raising=False)
    try:
        import gym  # noqa: F401
        pytest.skip("a real gym is installed")
    except ImportError:
        pass
    assert compat.register_gym_ids() is False  # nothing to register with
    calls = {}
    gym = types.ModuleType("gym")
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")
    reg.register = lambda id, entry_point=None, **kw: calls.__setitem__(id, entry_point)  # noqa: A002
    envs.registration, gym.envs = reg, envs
    for name, mod in (("gym", gym), ("gym.envs", envs), ("gym.envs.registration", reg)):
        monkeypatch.setitem(sys.modules, name, mod)
    assert compat.register_gym_ids() is True
    assert calls == {
        "gym_uav_collision_avoidance/UAVWorld2D-v0": "gym_uav_collision_avoidance_b200.compat:UAVWorld2D",
        "gym_uav_collision_avoidance/MultiUAVWorld2D-v0": "gym_uav_collision_avoidance_b200.compat:MultiUAVWorld2D",
    }
    # the entry points resolve to the drop-in classes with the reference's constructor signatures
    import importlib
    import inspect

    for entry in calls.values():
        mod, cls = entry.split(":")
        klass = getattr(importlib.import_module(mod), cls)
        assert "max_speed" in inspect.signature(klass.__init__).parameters


def test_nvtx_ranges_are_opt_in():
    """UAVCA_NVTX=1 wraps the batched entry points in named NVTX ranges at import (tracing); unset, the methods
    are the plain ones."""
    import subprocess
    import sys

    code = ("import gym_uav_collision_avoidance_b200.batched as b; "
            "print(hasattr(b.BatchedMultiUAVWorld2D.step, '__wrapped__'), hasattr(b.BatchedUAVWorld2D.rollout, '__wrapped__'))")
    env = dict(os.environ, PYTHONPATH=ROOT)
    on = subprocess.run([sys.executable, "-c", code], env=dict(env, UAVCA_NVTX="1"), capture_output=True, text=True, timeout=300)
    off = subprocess.run([sys.executable, "-c", code], env={k: v for k, v in env.items() if k != "UAVCA_NVTX"},
                         capture_output=True, text=True, timeout=300)
    assert on.stdout.split() == ["True", "True"], on.stderr[-500:]
    assert off.stdout.split() == ["False", "False"], off.stderr[-500:]
