"""Host-side noise plumbing of the planner (CPU tensors; no kernels): layouts, the pixel models' ShiftAug draw, and the
declared non-parity in-kernel-noise mode, which must not draw or store the two large tensors."""
import torch

from tdmpc2_b200.config import workload
from tdmpc2_b200.planner import Noise, alloc_noise, draw_noise


def test_iteration_major_layout_and_batched_draws():
    cfg = workload("tiny-mt", num_envs=3)
    H, N, P, A, I, K, E = cfg.horizon, cfg.num_samples, cfg.num_pi_trajs, cfg.action_dim, cfg.iterations, cfg.num_elites, 3
    nz = draw_noise(cfg, E, "cpu", generator=torch.Generator().manual_seed(1))
    assert nz.prior.shape == (E, H, P, A) and nz.r.shape == (I, E, H, N - P, A) and nz.pi.shape == (I, E, N, A)
    assert nz.qidx.shape == (I, E, 2) and nz.qidx.dtype == torch.int32 and nz.expo.shape == (E, K) and nz.final.shape == (E, A)
    assert all(t.is_contiguous() for t in nz.tensors()) and nz.shift is None
    assert bool((nz.qidx[..., 0] != nz.qidx[..., 1]).all()) and int(nz.qidx.max()) < cfg.num_q      # randperm(num_q)[:2]
    assert draw_noise(cfg, E, "cpu", eval_mode=True, generator=torch.Generator().manual_seed(1)).final is None
    # every CEM iteration reads one contiguous slab
    assert nz.r[1].is_contiguous() and nz.pi[2].is_contiguous()


def test_reference_order_single_env_consumes_the_generator_like_the_reference():
    cfg = workload("tiny", num_envs=1)
    H, N, P, A, I, K = cfg.horizon, cfg.num_samples, cfg.num_pi_trajs, cfg.action_dim, cfg.iterations, cfg.num_elites
    nz = draw_noise(cfg, 1, "cpu", generator=torch.Generator().manual_seed(5))
    g = torch.Generator().manual_seed(5)
    for t in range(H):
        assert torch.equal(nz.prior[0, t], torch.randn(P, A, generator=g))
    for it in range(I):
        assert torch.equal(nz.r[it, 0], torch.randn(H, N - P, A, generator=g))
        assert torch.equal(nz.pi[it, 0], torch.randn(N, A, generator=g))
        assert torch.equal(nz.qidx[it, 0], torch.randperm(cfg.num_q, generator=g)[:2].to(torch.int32))
    assert torch.equal(nz.expo[0], torch.empty(K).exponential_(generator=g))
    assert torch.equal(nz.final[0], torch.randn(A, generator=g))


def test_pixel_models_draw_shiftaug_first():
    cfg = workload("tiny-rgb", num_envs=1)
    nz = draw_noise(cfg, 1, "cpu", generator=torch.Generator().manual_seed(9))
    g = torch.Generator().manual_seed(9)
    want = torch.randint(0, 7, (1, 2), dtype=torch.float32, generator=g)            # layers.py:55, before every other draw
    assert torch.equal(nz.shift, want) and bool(((nz.shift >= 0) & (nz.shift <= 6)).all())
    assert torch.equal(nz.prior[0, 0], torch.randn(cfg.num_pi_trajs, cfg.action_dim, generator=g))
    env_major = Noise.from_env_major(nz.prior, nz.r.transpose(0, 1), nz.pi.transpose(0, 1), nz.qidx.transpose(0, 1), nz.expo,
                                     nz.final, shift=nz.shift)
    assert torch.equal(env_major.r, nz.r) and torch.equal(env_major.shift, nz.shift)


def test_in_kernel_noise_mode_allocates_no_large_tensors():
    cfg = workload("c2", rng="philox")
    nz = alloc_noise(cfg, cfg.num_envs, "cpu")
    assert nz.r is None and nz.pi is None
    assert sum(t.numel() for t in nz.tensors()) < 1_000_000                          # vs 115 M floats with torch's draws
    drawn = draw_noise(cfg, cfg.num_envs, "cpu", generator=torch.Generator().manual_seed(2), out=nz)
    assert drawn.r is None and float(drawn.prior.std()) > 0.9 and float(drawn.expo.min()) > 0
