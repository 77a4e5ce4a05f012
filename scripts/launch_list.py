"""Every kernel of ONE steady-state plan() of a bench workload, for the ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
        python scripts/launch_list.py c2

The step runs the eager launch chain (the CUDA graph bench.py replays holds the same kernels; ncu would need
--graph-profiling node to see inside it), noise draws included; the profiler range covers exactly one step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import bench_cfg
from tdmpc2_b200.synth import synth_state_dict
from tdmpc2_b200.tdmpc2 import TDMPC2

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = bench_cfg(wl, int(sys.argv[2]) if len(sys.argv) > 2 else None)
cfg.cuda_graph = False
dev = torch.device("cuda:0")
agent = TDMPC2(cfg, device=dev)
agent.load(synth_state_dict(cfg, seed=1))
E = cfg.num_envs
obs = torch.randn(E, cfg.obs_shape["state"][0], device=dev)
task = (torch.arange(E) % len(cfg.tasks)).to(torch.int32).to(dev) if cfg.multitask else None
agent._plan(obs, t0=True, task=task)
agent._plan(obs, t0=False, task=task)
torch.cuda.synchronize()
torch.cuda.profiler.start()
agent._plan(obs, t0=False, task=task)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one plan() step profiled")
