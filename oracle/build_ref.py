"""oracle/_ref: a travelling copy of the reference's environment core.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python; `/root/reference` exists only in the build container.  So that the GPU box can time
the LITERAL reference step next to the CUDA path (bench.py `cpu_baseline_literal`, `--workload c1`), this recipe
copies the five files of the env core

    gym_uav_collision_avoidance/__init__.py
    gym_uav_collision_avoidance/envs/{__init__,uav_agent,uav_world_2d,multi_uav_world_2d}.py

unmodified from where they lie under `/root/reference` into `oracle/_ref/` (git-ignored: reference sources never
enter this repository's history; not gpurun-ignored: the directory travels with the snapshot like a built .so).
`oracle/ref_loader.py` imports from `/root/reference` when it exists and from `oracle/_ref` otherwise, always through
the same `gym` / `pygame` / `turtle` stand-ins.

    python -m oracle.build_ref        # (re)creates oracle/_ref/ ; no-op where /root/reference is absent
"""
from __future__ import annotations

import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("UAVCA_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(_HERE, "_ref")
FILES = (
    "gym_uav_collision_avoidance/__init__.py",
    "gym_uav_collision_avoidance/envs/__init__.py",
    "gym_uav_collision_avoidance/envs/uav_agent.py",
    "gym_uav_collision_avoidance/envs/uav_world_2d.py",
    "gym_uav_collision_avoidance/envs/multi_uav_world_2d.py",
)


def build_ref() -> str | None:
    """Copy the env core into oracle/_ref/.  Returns the directory, or None where the reference is absent (then an
    earlier copy, if any, is left alone)."""
    if not os.path.isdir(os.path.join(REF_SRC, "gym_uav_collision_avoidance", "envs")):
        return REF_DST if os.path.isdir(REF_DST) else None
    for rel in FILES:
        dst = os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, rel), dst)
    with open(os.path.join(REF_DST, "README"), "w") as f:
        f.write("Unmodified copies of the reference's env core, made by oracle/build_ref.py.  Git-ignored on purpose.\n")
    return REF_DST


if __name__ == "__main__":
    print(build_ref())
