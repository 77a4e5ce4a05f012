"""scripts/torch_gpu_baseline.py (the batched eager-PyTorch baseline of SURVEY.md 8(d)(iii)) computes the same plan as
the oracle: checked here on CPU so that its GPU timing next to the fused kernels is a timing of the right algorithm."""
import importlib.util
import os

import pytest
import torch

from oracle.plan_oracle import balance_termination, draw_noise, plan_oracle
from tdmpc2_b200.config import workload
from tdmpc2_b200.synth import synth_state_dict
from helpers import stable_positions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("torch_gpu_baseline", os.path.join(ROOT, "scripts", "torch_gpu_baseline.py"))
tgb = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tgb)


@pytest.mark.parametrize("wl,over", [("tiny", {}), ("tiny-mt", {}), ("tiny", {"episodic": True})])
def test_batched_torch_planner_matches_oracle(wl, over):
    cfg = workload(wl, **over)
    E = cfg.num_envs
    sd = synth_state_dict(cfg, seed=17, perturb=True)
    if cfg.episodic:
        balance_termination(cfg, sd)
    g = torch.Generator().manual_seed(4)
    obs = torch.randn(E, cfg.obs_shape["state"][0], generator=g)
    prev = 0.3 * torch.randn(E, cfg.horizon, cfg.action_dim, generator=g)
    t0 = [bool(i % 2) for i in range(E)]
    task = [(i + 1) % len(cfg.tasks) for i in range(E)] if cfg.multitask else None
    noise = draw_noise(cfg, 50, E)
    want = plan_oracle(cfg, sd, obs, task=task, t0=t0, prev_mean=prev, noise=noise)
    pl = tgb.BatchedTorchPlanner(cfg, sd, "cpu")
    a, mean, values = pl.plan(obs, None if task is None else torch.tensor(task), torch.tensor(t0), prev, noise)
    ok = torch.ones_like(want.values, dtype=torch.bool)
    if cfg.episodic:
        ok = want.term_margin > 1e-5
    # first iteration: same inputs -> same values up to re-association (batched matmuls vs per-env F.linear)
    assert float((values[:, 0] - want.values[:, 0]).abs()[ok[:, 0]].max()) < 2e-5
    if bool(ok.all()) and all(bool(stable_positions(want.values[e, it], cfg.num_elites, 1e-4).all())
                              for e in range(E) for it in range(cfg.iterations)):
        assert torch.allclose(mean, want.mean, atol=1e-4, rtol=0)
        assert torch.allclose(a, want.action, atol=1e-4, rtol=0)
