"""GPU: K steps per launch (`uavca_rollout`) against K single steps and against the oracle; the device-side replay head.

`rollout` runs the very same step code with the env state held in registers, so it must be BIT-identical to K calls of
`step()` fed the same actions — state blob, observations, rewards, done flags, reset masks, counters."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

from _golden import obs_close

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


def _b200():
    import gym_uav_collision_avoidance_b200 as G

    return G


def _pair(kind, B, N, **kw):
    G = _b200()
    if kind == "single":
        mk = lambda: G.BatchedUAVWorld2D(B, reset_mode=O.RESET_ON_ANY_DONE, **kw)  # noqa: E731
    else:
        mk = lambda: G.BatchedMultiUAVWorld2D(B, num_agents=N, reset_mode=O.RESET_ON_DONE0, **kw)  # noqa: E731
    a, b = mk(), mk()
    a.reset()
    b.reset()
    return a, b


@pytest.mark.parametrize("n,B", [(2, 3000), (5, 1026), (7, 2050), (8, 4099), (10, 1500), (16, 1001), (32, 515), (24, 300),
                                 (17, 131), (20, 262), (25, 77), (28, 90), (11, 190), (12, 170)])
def test_rollout_with_action_block_equals_single_steps(n, B):
    if (B * n * 10 * 4) % 16:
        B += 1  # K > 1 needs 16-byte aligned step blocks
    K = 37
    e1, e2 = _pair("multi", B, n, seed=50 + n, max_episode_steps=15, x_size=20.0, y_size=20.0)
    gen = torch.Generator(device="cuda").manual_seed(n)
    acts = torch.rand((K, B, n, 2), generator=gen, device="cuda") * 20 - 10
    out = e1.rollout(K, acts, want_final_obs=True, want_reset_mask=True)
    e2.enable_final_obs()
    resets = 0
    for k in range(K):
        o, r, d, info = e2.step(acts[k])
        assert torch.equal(out["obs"][k], o), f"obs differs at step {k}"
        assert torch.equal(out["reward"][k], r) and torch.equal(out["done"][k], d), f"reward/done differ at step {k}"
        assert torch.equal(out["reset_mask"][k], info["reset_mask"]) and torch.equal(out["final_obs"][k], info["final_obs"])
        resets += int(info["reset_mask"].sum())
    assert resets > B  # every env went through the step limit at least twice
    assert torch.equal(e1.state.blob, e2.state.blob), "state after the rollout differs from K single steps"
    assert torch.equal(e1.obs, e2.obs) and torch.equal(e1.reward, e2.reward) and torch.equal(e1.done, e2.done)


@pytest.mark.parametrize("mode", ["cartesian", "polar", "scaled"])
@pytest.mark.parametrize("n", [3, 8, 20, 32])
def test_rollout_with_philox_actions(n, mode):
    """On-device actions: the draws equal the oracle's restatement of the stream, the rollout equals K single steps fed
    `sample_actions`, and both agree with the oracle (flags bit-exact, outputs within tolerance)."""
    B, K, seed, step0 = 512, 40, 0xAC7 + n, 1001  # odd step0: the first step uses the second half of a Philox block
    kw = dict(seed=9 + n, max_episode_steps=25)
    e1, e2 = _pair("multi", B, n, **kw)
    cfg = O.multi_config(B, n, reset_mode=O.RESET_ON_DONE0, **kw)
    orc = O.Oracle(cfg, nthreads=4)
    orc.reset()
    out = e1.rollout(K, None, action_mode=mode, action_seed=seed, step0=step0, want_actions=True, want_reset_mask=True)
    single_mode = "scaled" if mode == "cartesian" else mode  # policy-space draws: "cartesian" spans the action box
    omode = {"polar": O.ACTION_POLAR, "scaled": O.ACTION_SCALED}[single_mode]
    for k in range(K):
        a = e2.sample_actions(step0 + k, seed)
        assert torch.equal(out["actions"][k], a)
        assert np.array_equal(a.cpu().numpy(), orc.sample_actions(step0 + k, seed)), "Philox action stream differs from the oracle"
        o, r, d, info = e2.step(a, action_mode=single_mode)
        assert torch.equal(out["obs"][k], o) and torch.equal(out["reward"][k], r) and torch.equal(out["done"][k], d)
        if mode == "polar":  # the device maps polar actions through float32 sin/cos: compared with the oracle elsewhere
            continue
        ref = orc.step(a.cpu().numpy(), action_mode=omode)
        assert np.array_equal(d.cpu().numpy(), ref["done"]) and np.array_equal(out["reset_mask"][k].cpu().numpy(), ref["reset_mask"])
        assert np.array_equal(e2.state.pos.cpu().numpy(), orc.state.pos) and np.array_equal(e2.state.vel.cpu().numpy(), orc.state.vel)
        rr = r.cpu().numpy()
        assert (np.abs(rr - ref["reward"]) <= RTOL * np.abs(ref["reward"]) + ATOL).all()
        assert obs_close(o.cpu().numpy(), ref["obs"], RTOL, ATOL).all()
    assert torch.equal(e1.state.blob, e2.state.blob)
    a = out["actions"]
    assert float(a.min()) >= -1.0 and float(a.max()) < 1.0 and abs(float(a.mean())) < 0.01


@pytest.mark.parametrize("n,circular", [(40, False), (6, True), (33, False)])
def test_rollout_on_the_general_kernel(n, circular):
    """Envs wider than a warp and the float64 world: uavca_rollout is the same call (K launches inside), bit-identical to K
    single steps — with an action block and with the Philox stream."""
    G = _b200()
    B, K = 64, 12
    kw = dict(num_agents=n, reset_mode=O.RESET_ON_DONE0, max_episode_steps=9, seed=70 + n, circular=circular)
    e1, e2, e3, e4 = (G.BatchedMultiUAVWorld2D(B, **kw) for _ in range(4))
    for e in (e1, e2, e3, e4):
        e.reset()
    gen = torch.Generator(device="cuda").manual_seed(n)
    acts = torch.rand((K, B, n, 2), generator=gen, device="cuda") * 20 - 10
    out = e1.rollout(K, acts, want_final_obs=True, want_reset_mask=True)
    e2.enable_final_obs()
    for k in range(K):
        o, r, d, info = e2.step(acts[k])
        assert torch.equal(out["obs"][k], o) and torch.equal(out["reward"][k], r) and torch.equal(out["done"][k], d), k
        assert torch.equal(out["reset_mask"][k], info["reset_mask"]) and torch.equal(out["final_obs"][k], info["final_obs"])
    assert torch.equal(e1.state.blob, e2.state.blob) and int(out["reset_mask"].sum()) >= B
    out = e3.rollout(K, None, action_mode="polar", action_seed=5, step0=3)  # no action_out: the handle's scratch
    for k in range(K):
        o, r, d, _ = e4.step(e4.sample_actions(3 + k, 5), action_mode="polar")
        assert torch.equal(out["obs"][k], o) and torch.equal(out["reward"][k], r) and torch.equal(out["done"][k], d), k
    assert torch.equal(e3.state.blob, e4.state.blob)


@pytest.mark.parametrize("f32", [False, True])
def test_rollout_single_world(f32):
    B, K = 4097, 60
    e1, e2 = _pair("single", B, 1, seed=3, max_episode_steps=40, float32_first_step=f32)
    gen = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand((K, B, 1, 2), generator=gen, device="cuda") * 24 - 12
    out = e1.rollout(K, acts, want_final_obs=True, want_reset_mask=True)
    e2.enable_final_obs()
    for k in range(K):
        o, r, d, info = e2.step(acts[k])
        assert torch.equal(out["obs"][k], o) and torch.equal(out["reward"][k], r) and torch.equal(out["done"][k], d)
        assert torch.equal(out["distance"][k], info["distance"]) and torch.equal(out["reset_mask"][k], info["reset_mask"])
        assert torch.equal(out["final_obs"][k], info["final_obs"])
    assert torch.equal(e1.state.blob, e2.state.blob)
    # Philox actions: the run.py loop
    out = e1.rollout(K, None, action_seed=11, step0=0, want_actions=True)
    for k in range(K):
        a = e2.sample_actions(k, 11)
        assert torch.equal(out["actions"][k], a)
        o, r, d, _ = e2.step(a, action_mode="scaled")
        assert torch.equal(out["obs"][k], o) and torch.equal(out["done"][k], d)
    assert torch.equal(e1.state.blob, e2.state.blob)


def test_rollout_replays_from_a_cuda_graph():
    """`out=` reuses the output blocks; with Philox actions the caller advances `step0` — here the graph holds two
    rollouts (even / odd chunks share buffers), and is compared with eager calls."""
    G = _b200()
    B, N, K = 2048, 8, 16
    kw = dict(num_agents=N, seed=4, reset_mode=O.RESET_ON_DONE0, max_episode_steps=50)
    eager, graphed = G.BatchedMultiUAVWorld2D(B, **kw), G.BatchedMultiUAVWorld2D(B, **kw)
    eager.reset()
    graphed.reset()
    acts = torch.zeros((K, B, N, 2), device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        buf = graphed.rollout(K, acts)
    torch.cuda.synchronize()
    graphed.state.blob.copy_(eager.state.blob)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        graphed.rollout(K, acts, out=buf)
    gen = torch.Generator(device="cuda").manual_seed(2)
    for _ in range(6):
        acts.copy_(torch.rand((K, B, N, 2), generator=gen, device="cuda") * 20 - 10)
        g.replay()
        ref = eager.rollout(K, acts)
        assert torch.equal(buf["obs"], ref["obs"]) and torch.equal(buf["done"], ref["done"])
    assert torch.equal(eager.state.blob, graphed.state.blob) and torch.equal(eager.obs, graphed.obs)


def test_rollout_rejects_bad_arguments():
    G = _b200()
    env = G.BatchedMultiUAVWorld2D(33, num_agents=5)  # 33*5*40 bytes is not a multiple of 16
    env.reset()
    with pytest.raises(G.UavcaError):
        env.rollout(4, None)
    env.rollout(1, None)  # a single step has no alignment requirement
    with pytest.raises(ValueError):
        env.rollout(3, torch.zeros((2, 33, 5, 2), device="cuda"))


def test_replay_ring_head_lives_on_the_device():
    """A CUDA-graph replay of the append must continue where the previous replay stopped (the head used to be a host
    integer baked into the captured launch)."""
    G = _b200()
    from gym_uav_collision_avoidance_b200.replay import DeviceReplay

    M, cap = 96, 1000
    rb = DeviceReplay(cap, 10, 2)
    obs = torch.zeros((M, 10), device="cuda")
    act = torch.zeros((M, 2), device="cuda")
    rew = torch.zeros(M, device="cuda")
    nxt = torch.zeros((M, 10), device="cuda")
    done = torch.zeros(M, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        rb.push(obs, act, rew, nxt, done)
    torch.cuda.synchronize()
    rb.meta.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rb.push(obs, act, rew, nxt, done)
    ref_r = torch.zeros(cap, device="cuda")
    ref_m = torch.zeros(cap, device="cuda")
    pos = 0
    for it in range(25):  # 2,400 transitions through a ring of 1,000: wraps twice
        rew.copy_(torch.arange(M, device="cuda") + 1000.0 * it)
        done.copy_((torch.arange(M, device="cuda") + it) % 3 == 0)
        g.replay()
        idx = (pos + torch.arange(M, device="cuda")) % cap
        ref_r[idx] = rew
        ref_m[idx] = 1.0 - done.float()
        pos = (pos + M) % cap
        assert rb.position == pos and len(rb) == min(cap, (it + 1) * M)
    assert torch.equal(rb.reward, ref_r) and torch.equal(rb.mask, ref_m)
