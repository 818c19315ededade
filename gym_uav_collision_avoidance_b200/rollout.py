"""Batched acting path around the env step ("next" rows of SURVEY.md §8f): one policy forward over all B*N UAVs,
the polar action map fused into the step kernel, transitions appended to the device replay ring — nothing leaves
the GPU between `obs` and the next `obs`.

The reference acts one UAV at a time: `SAC.select_action` copies a (10,) observation to the GPU, samples, copies the
action back (pytorch_sac_temp/sac.py:38-44), N times per env step (test_sac_multi.py:69-80), then maps the action
to cartesian on the host (:77-80) and pushes N tuples into a Python list (:101-103).  All UAVs share one policy
(:90-91), so the batched equivalent is a single [B*N, 10] forward.  The networks are stock PyTorch (cuBLAS): dense
MLPs are the learner's side of the boundary, not part of the hot path this package accelerates.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .replay import DeviceReplay

LOG_SIG_MAX, LOG_SIG_MIN, EPS = 2, -20, 1e-6  # pytorch_sac_temp/model.py:6-8


class GaussianPolicy(nn.Module):
    """Same architecture and sampling semantics as the reference policy (pytorch_sac_temp/model.py:64-101):
    10 -> 256 -> 256 -> (mean, log_std), tanh-squashed Gaussian.  Random-init here (the reference ships no weights);
    `load_state_dict` accepts a reference checkpoint's `policy_state_dict` (same parameter names)."""

    def __init__(self, num_inputs=10, num_actions=2, hidden=256):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden)
        self.linear2 = nn.Linear(hidden, hidden)
        self.mean_linear = nn.Linear(hidden, num_actions)
        self.log_std_linear = nn.Linear(hidden, num_actions)
        for m in self.modules():  # weights_init_ (model.py:11-14)
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=1)
                nn.init.constant_(m.bias, 0)

    def forward(self, state):
        x = F.relu(self.linear1(state))
        x = F.relu(self.linear2(x))
        return self.mean_linear(x), torch.clamp(self.log_std_linear(x), min=LOG_SIG_MIN, max=LOG_SIG_MAX)

    def sample(self, state):
        """-> (action, log_prob, eval_action) as model.py:86-101.  Note the reference's third output is
        tanh(normal.sample()), not tanh(mean) (kept as is)."""
        mean, log_std = self.forward(state)
        std = log_std.exp()
        x_t = mean + std * torch.randn_like(mean)
        y_t = torch.tanh(x_t)
        log_prob = -((x_t - mean) ** 2) / (2 * std * std) - log_std - 0.5 * math.log(2 * math.pi)
        log_prob = (log_prob - torch.log(1 - y_t.pow(2) + EPS)).sum(1, keepdim=True)
        return y_t, log_prob, torch.tanh(mean + std * torch.randn_like(mean))

    @torch.no_grad()
    def act(self, state, evaluate=False):
        a, _, e = self.sample(state)
        return e if evaluate else a


class BatchedRollout:
    """obs -> policy -> step(action_mode) -> replay, all on the device; optionally replayed from a CUDA graph."""

    def __init__(self, env, policy: Optional[nn.Module] = None, replay: Optional[DeviceReplay] = None,
                 action_mode="polar", evaluate=False, warmup_uniform=False):
        self.env, self.policy, self.replay = env, policy, replay
        self.action_mode, self.evaluate, self.warmup_uniform = action_mode, evaluate, warmup_uniform
        B, N, D = env.num_envs, env.num_agents, env.obs_dim
        self.state = torch.zeros((B, N, D), dtype=torch.float32, device=env.device)  # observation the action was taken on
        self.action = torch.zeros((B, N, 2), dtype=torch.float32, device=env.device)
        if replay is not None:
            env.enable_final_obs()
        self._graph = None
        self.steps = 0

    def reset(self):
        self.env.reset()
        return self.env.obs

    def _act(self):
        env = self.env
        self.state.copy_(env.obs)
        if self.policy is None or self.warmup_uniform:  # test_sac_multi.py:72-73: uniform actions during warm-up
            self.action.uniform_(-1.0, 1.0)
        else:
            a = self.policy.act(self.state.view(-1, env.obs_dim), evaluate=self.evaluate)
            self.action.copy_(a.view_as(self.action))

    def _env_step(self):
        env = self.env
        if env.num_agents == 1 and env.obs_dim == 4:
            env.step(self.action, action_mode=self.action_mode)
        else:
            env.step(self.action, evaluate=self.evaluate, action_mode=self.action_mode)

    def step(self):
        self._act()
        self._env_step()
        if self.replay is not None:
            # next_state = the step's own observation (before any auto-reset), mask = float(not done): test_sac_multi.py:101-103
            self.replay.push(self.state, self.action, self.env.reward, self.env.final_obs, self.env.done)
        self.steps += 1

    def run(self, steps: int):
        for _ in range(steps):
            self.step()

    def success_collision_rates(self):
        """SR / CR over the finished episodes (test_sac_multi.py:164-179): counts / (N * episodes)."""
        s = self.env.stats()
        denom = max(1, s["episodes"]) * self.env.num_agents
        return s["reach"] / denom, s["collisions"] / denom, s["episodes"]
