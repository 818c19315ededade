"""CPU: oracle/uav_oracle.c replayed against the golden vectors the literal reference produced
(oracle/gen_golden.py).  Everything is compared for EXACT equality — flags, float32 positions, float64
velocities, float64 rewards and float64 observations."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_loader as R

from _golden import Case, golden_names


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    case = Case(name)
    orc = O.Oracle(case.config())
    orc.state = case.init_state()
    orc.set_pool(case.pool_state())
    z = case.z
    assert np.array_equal(orc.observe(), z["obs0"]), "reset observation"
    for t in range(case.T):
        out = orc.step(z["action"][t], evaluate=case.evaluate, want_final_obs=True)
        assert np.array_equal(out["done"], z["done"][t]), f"done flags, step {t}"
        assert np.array_equal(out["reset_mask"], z["reset_mask"][t]), f"reset mask, step {t}"
        assert np.array_equal(out["reward"], z["reward"][t]), f"reward, step {t}"
        assert np.array_equal(out["final_obs"], case.final_obs(t)), f"terminal observation, step {t}"
        assert np.array_equal(out["obs"], z["obs"][t]), f"observation, step {t}"
        assert np.array_equal(orc.state.pos, z["pos"][t]), f"position, step {t}"
        assert np.array_equal(orc.state.vel, z["vel"][t]), f"velocity, step {t}"
        assert np.array_equal(orc.state.prev, z["prev"][t]), f"prev_distance, step {t}"
        assert np.array_equal(orc.state.steps, z["steps"][t]), f"env.steps, step {t}"
        if case.kind == "multi":
            assert np.array_equal(orc.state.flags, z["flags"][t]), f"parked/collided latches, step {t}"
            assert np.array_equal(orc.state.reach, z["reach"][t]), f"target_reach_count, step {t}"
            assert np.array_equal(orc.state.coll, z["coll"][t]), f"collision_count, step {t}"
        else:
            assert np.array_equal(out["distance"], z["distance"][t]), f"info distance, step {t}"


@pytest.mark.parametrize("name", golden_names("circular"))
def test_oracle_matches_reference_golden_circular(name):
    """Episodes started by reset(circular=True): the reference keeps float64 locations (multi_uav_world_2d.py:157-163), the
    oracle's float64 world must reproduce them bit for bit — positions, velocities, rewards, observations, flags."""
    case = Case(name)
    orc = O.Oracle(case.config())
    z = case.z
    assert np.array_equal(orc.reset(), z["obs0"]), "reset(circular=True) observation"
    assert np.allclose(np.linalg.norm(orc.state.pos64, axis=-1), 20.0, atol=1e-12)
    for t in range(case.T):
        steps_before = orc.state.steps.copy()
        out = orc.step(z["action"][t], evaluate=case.evaluate, want_final_obs=True)
        assert np.array_equal(out["done"], z["done"][t]), f"done flags, step {t}"
        assert np.array_equal(out["reset_mask"], z["reset_mask"][t]), f"reset mask, step {t}"
        assert np.array_equal(out["reward"], z["reward"][t]), f"reward, step {t}"
        assert np.array_equal(out["final_obs"], z["final_obs"][t]), f"terminal observation, step {t}"
        assert np.array_equal(out["obs"], z["obs"][t]), f"observation, step {t}"
        assert np.array_equal(orc.state.pos64, z["pos64"][t]), f"position, step {t}"
        assert np.array_equal(orc.state.vel, z["vel"][t]), f"velocity, step {t}"
        assert np.array_equal(orc.state.prev64, z["prev64"][t]), f"prev_distance, step {t}"
        assert np.array_equal(orc.state.flags, z["flags"][t]), f"parked/collided latches, step {t}"
        keep = z["reset_mask"][t] == 0  # the golden counters were read before a restart zeroed them
        assert np.array_equal(orc.state.steps[keep], z["steps"][t][keep]) and np.array_equal(steps_before + 1, z["steps"][t])
        assert np.array_equal(orc.state.reach[keep], z["reach"][t][keep]) and np.array_equal(orc.state.coll[keep], z["coll"][t][keep])
        assert np.array_equal(orc.state.pos, z["pos64"][t].astype(np.float32))  # the float32 mirrors follow


def test_golden_cases_exercise_the_interesting_events():
    ev = dict(done=0, resets=0, reach=0, coll=0, parked=0)
    for name in golden_names("multi"):
        z = Case(name).z
        ev["done"] += int(z["done"].sum())
        ev["resets"] += int(z["reset_mask"].sum())
        ev["reach"] += int(z["reach"].max())
        ev["coll"] += int(z["coll"].max())
        ev["parked"] += int((z["flags"] & 1).sum())
    assert ev["done"] > 1000 and ev["resets"] > 10 and ev["reach"] > 10 and ev["coll"] > 10 and ev["parked"] > 1000


@pytest.mark.skipif(not R.reference_available(), reason="reference checkout only exists in the build container")
@pytest.mark.parametrize("n", [3, 7])
def test_oracle_matches_live_reference(n):
    """Where the reference is present, regenerate a small case live (fresh seed) and compare."""
    from oracle import gen_golden as G

    case = dict(name="live", N=n, E=4, T=120, seed=900 + n, evaluate=0, reset_mode=O.RESET_ON_DONE0, max_steps=60)
    res = G.run_reference_multi(case)
    cfg = O.multi_config(4, n, reset_mode=O.RESET_ON_DONE0, max_episode_steps=60, reset_source=O.SOURCE_POOL)
    orc = O.Oracle(cfg)
    pool = O.State(G.POOL, n)
    for f in ("pos", "vel", "tgt", "init", "prev", "flags"):
        getattr(orc.state, f)[...] = res["init_" + f]
        getattr(pool, f)[...] = res["pool_" + f]
    orc.state.episode[...] = 1
    orc.set_pool(pool)
    for t in range(case["T"]):
        out = orc.step(res["action"][t], want_final_obs=True)
        assert np.array_equal(out["done"], res["done"][t])
        assert np.array_equal(out["reward"], res["reward"][t])
        assert np.array_equal(out["obs"], res["obs"][t])
        assert np.array_equal(out["final_obs"], res["final_obs"][t])
        assert np.array_equal(orc.state.pos, res["pos"][t])
        assert np.array_equal(orc.state.vel, res["vel"][t])
