"""Helpers shared by the CPU (oracle) and GPU (CUDA) parity tests: load a golden case written by
oracle/gen_golden.py from the literal reference and turn it into oracle-side config/state objects."""
import glob
import json
import os

import numpy as np

from oracle import oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(kind=None):
    """Golden cases by name prefix; without a prefix: the float32-world cases (multi_*, single_*), which share one replay
    loop — the float64-world cases (circular_*) have their own tests."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if kind:
        return [n for n in names if n.startswith(kind)]
    return [n for n in names if not n.startswith("circular")]


class Case:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.kind = self.meta["kind"]
        self.E, self.T = self.meta["E"], self.meta["T"]
        self.N = self.meta.get("N", 1)
        self.evaluate = bool(self.meta.get("evaluate", 0))

    def config(self) -> O.Config:
        m = self.meta
        if self.kind == "single":
            return O.single_config(self.E, reset_mode=m["reset_mode"], reset_source=O.SOURCE_POOL,
                                   single_f32_first_step=m["f32"])
        if self.kind == "circular":  # reset(circular=True): the float64 world, every episode restarts on the ring
            return O.multi_config(self.E, self.N, reset_mode=m["reset_mode"], max_episode_steps=m["max_steps"], circular=1)
        return O.multi_config(self.E, self.N, reset_mode=m["reset_mode"], max_episode_steps=m["max_steps"],
                              reset_source=O.SOURCE_POOL)

    def _state(self, prefix, n_envs) -> O.State:
        st = O.State(n_envs, self.N)
        for f in ("pos", "vel", "tgt", "init", "prev", "flags"):
            getattr(st, f)[...] = self.z[prefix + f]
        st.episode[...] = 1  # the golden harness starts after an initial reset()
        return st

    def init_state(self) -> O.State:
        return self._state("init_", self.E)

    def pool_state(self) -> O.State:
        return self._state("pool_", self.meta["pool"])

    def final_obs(self, t):
        return self.z["final_obs"][t] if "final_obs" in self.z.files else self.z["obs"][t]


def circ_close(a, b, rtol=1e-5, atol=1e-6):
    """|a-b| <= rtol*|b| + atol for angle features normalised by pi, compared on the circle (-1 == +1)."""
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    d = np.minimum(d, 2.0 - d)
    return d <= rtol * np.abs(b) + atol


ANGLE_FEATURES_MULTI = (1, 3, 5, 6, 8, 9)
ANGLE_FEATURES_SINGLE = (1, 3)


def obs_close(obs, ref, rtol=1e-5, atol=1e-6):
    """Elementwise parity mask for an observation tensor [..., D] against the float64 reference."""
    D = ref.shape[-1]
    ang = ANGLE_FEATURES_SINGLE if D == 4 else ANGLE_FEATURES_MULTI
    obs = np.asarray(obs, np.float64)
    ok = np.abs(obs - ref) <= rtol * np.abs(ref) + atol
    for k in ang:
        ok[..., k] = circ_close(obs[..., k], ref[..., k], rtol, atol)
    return ok
