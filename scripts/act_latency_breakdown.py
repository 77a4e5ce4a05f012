"""Where one-environment act() time goes (c1: dog-run 5M, E = 1): wall clock per act(), GPU time of the graph replay alone
(CUDA events), host time of the noise draws (reference order vs one batched draw), H2D / D2H."""
import statistics
import sys
import time

import torch

sys.path.insert(0, ".")
from tdmpc2_b200.config import workload
from tdmpc2_b200.planner import draw_noise
from tdmpc2_b200.synth import synth_state_dict
from tdmpc2_b200.tdmpc2 import TDMPC2


def med(f, n=30):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t))
    return statistics.median(ts)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c1"
    cfg = workload(wl, num_envs=1, e1_interleaved=False)
    sd = synth_state_dict(cfg, seed=1)
    obs = torch.randn(cfg.obs_shape["state"][0]).pin_memory()
    task = 0 if cfg.multitask else None

    def warm(c):
        a = TDMPC2(c, device="cuda:0")
        a.load(sd)
        a.act(obs, t0=True, task=task)
        for _ in range(5):
            a.act(obs, t0=False, task=task)
        return a

    inter = warm(workload(wl, num_envs=1))                      # default: draws interleaved with the launches
    print(f"workload {wl}  engine {inter.planner.iter_engine}  iterations {inter.cfg.iterations}")
    print(f"act() wall, interleaved draws   {med(lambda: inter.act(obs, t0=False, task=task)):.3f} ms")
    del inter
    agent = warm(cfg)                                           # all draws, then one graph replay
    pl = agent.planner
    st = pl._graphs[False]
    print(f"act() wall, draws then graph    {med(lambda: agent.act(obs, t0=False, task=task)):.3f} ms   ({st['launches']} launches per graph)")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gts = []
    for _ in range(30):
        torch.cuda.synchronize()
        e0.record(); st["graph"].replay(); e1.record()
        torch.cuda.synchronize()
        gts.append(e0.elapsed_time(e1))
    print(f"graph replay, GPU (events)      {statistics.median(gts):.3f} ms")
    print(f"graph replay, wall              {med(lambda: st['graph'].replay()):.3f} ms")
    print(f"noise draws, reference order    {med(lambda: draw_noise(cfg, 1, 'cuda:0', out=st['noise'])):.3f} ms")
    print(f"noise draws, one batched draw   {med(lambda: draw_noise(cfg, 1, 'cuda:0', out=st['noise'], reference_order=False)):.3f} ms")
    dobs = torch.empty_like(obs, device="cuda:0")
    print(f"obs H2D + action D2H            {med(lambda: (dobs.copy_(obs, non_blocking=True), st['action'].cpu())):.3f} ms")
    # per-kernel GPU time of the chain, launched eagerly
    from tdmpc2_b200.planner import Noise
    nz = st["noise"]
    evs = []
    def stamp():
        e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
    stamp(); pl.prologue(st["obs"], st["task"], st["t0"], st["prev"], nz.prior); stamp()
    for it in range(cfg.iterations):
        pl.iterate(nz.r[it], nz.pi[it], nz.qidx[it]); stamp()
    pl.epilogue(nz.expo, nz.final, st["action"], st["new_mean"]); stamp()
    torch.cuda.synchronize()
    d = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]
    print("eager chain, GPU ms: prologue %.3f | iterations %s | epilogue %.3f" % (d[0], " ".join(f"{x:.3f}" for x in d[1:-1]), d[-1]))


if __name__ == "__main__":
    main()
