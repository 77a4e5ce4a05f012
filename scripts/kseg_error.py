"""Accuracy / speed of the wide layers' K-segmented accumulation (tdmpc2_planner_set_kseg): trajectory-value error of
one environment of a wide preset against the CPU oracle, and the CEM-iteration time, per segment length.

    python scripts/kseg_error.py c4 [E] [ksegs...]      (GPU box; the oracle runs once on the host)
"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
torch.set_num_threads(min(16, os.cpu_count() or 1))
from tdmpc2_b200.config import workload
from tdmpc2_b200.synth import synth_state_dict
from tdmpc2_b200.planner import Planner
from oracle.plan_oracle import plan_oracle
from helpers import mixed_noise

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ksegs = [int(x) for x in sys.argv[3:]] or [0, 2048, 1024, 512, 256]
cfg = workload(wl, num_envs=E)
sd = synth_state_dict(cfg, seed=9)
g = torch.Generator().manual_seed(33)
obs = torch.randn(E, cfg.obs_shape["state"][0], generator=g)
prev = 0.3 * torch.randn(E, cfg.horizon, cfg.action_dim, generator=g)
t0 = torch.zeros(E, dtype=torch.uint8)
task = (torch.arange(E) * 7 + 3) % len(cfg.tasks) if cfg.multitask else None
env = [1 % E]
nz, on = mixed_noise(cfg, E, env, 400)
t = time.perf_counter()
want = plan_oracle(cfg, sd, obs[env], task=None if task is None else [int(task[e]) for e in env], t0=[False], prev_mean=prev[env], noise=on)
print(f"oracle: {time.perf_counter() - t:.1f} s", flush=True)
taskv = None if task is None else task.to(torch.int32).cuda()
for ks in ksegs:
    os.environ["TDMPC2_B200_KSEG"] = str(ks)
    pl = Planner(cfg, E, "cuda:0")
    pl.pack(sd)
    a, m, tr = pl.plan(obs.cuda(), taskv, t0.cuda(), prev.cuda(), nz, trace=True)
    torch.cuda.synchronize()
    errs = [float((tr["values"][env[0], it].cpu() - want.values[0, it]).abs().max()) for it in range(cfg.iterations)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pl.iterate(nz.r[0], nz.pi[0], nz.qidx[0]); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"workload": wl, "E": E, "kseg": ks, "max_value_err_per_iter": [f"{x:.2e}" for x in errs],
                      "mean_err": float((m[env[0]].cpu() - want.mean[0]).abs().max()), "iter_ms": e0.elapsed_time(e1)}), flush=True)
    del pl
