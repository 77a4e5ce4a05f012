#!/usr/bin/env python
"""Benchmark of the TD-MPC2 planning hot path on B200 (contract: see DESIGN.md 'Measurement').

    python bench.py --gpus 1 --steps 10 --warmup 3            # this build
    python bench.py --impl reference --steps 5 --warmup 3     # reference algorithm on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full plan() over the batch of environments: noise draws,
prologue (encode + policy-prior rollouts), I CEM iterations, epilogue -- and the
action all-gather when the environment axis is sharded (N > 1).
Metric (BASELINE.json): planning steps/sec = E * num_samples * horizon / t_plan.
Workload: BASELINE.json configs[1] ("c2": dog-run 5M model, 256 envs per GPU,
num_samples 512, horizon 3, iterations 6), synthetic weights/observations.
Weak scaling: every rank plans its own 256 environments.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tdmpc2_b200.config import workload, flops_per_env      # noqa: E402
from tdmpc2_b200.synth import synth_state_dict              # noqa: E402

METRIC = "planning steps/sec (num_envs x num_samples x horizon per plan() call)"
UNIT = "steps/s"
WORKLOAD = "c2"
CPU_THREADS = int(os.environ.get("TDMPC2_CPU_THREADS", "16"))


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_run(cfg, envs_per_step: int, steps: int, warmup: int, budget_s: float = 120.0):
    """The reference algorithm (oracle port: fp32 PyTorch on CPU, eager, all host threads)
    on the same workload.  The reference has no environment axis: environments are planned
    one after another, exactly what evaluate.py's loop would do."""
    from oracle.plan_oracle import OracleModel, draw_noise, plan_oracle
    # Intra-op threads: the reference's GEMMs are [512 x 550] x [550 x 512]; eager PyTorch stops
    # scaling (and collapses from barrier overhead) far below a 100+ core host.  Use what it can use.
    cores = min(os.cpu_count() or 1, CPU_THREADS)
    torch.set_num_threads(cores)
    sd = synth_state_dict(cfg, seed=1)
    model = OracleModel(cfg, sd)
    g = torch.Generator().manual_seed(2)
    obs = torch.randn(envs_per_step, cfg.obs_shape["state"][0], generator=g)
    prev = torch.zeros(envs_per_step, cfg.horizon, cfg.action_dim)
    times = []
    t_begin = time.perf_counter()
    for s in range(warmup + steps):
        if times and time.perf_counter() - t_begin > budget_s:
            break                                    # bounded sample: never let the CPU leg run away
        noise = draw_noise(cfg, 3 + 1000 * s, envs_per_step)
        t = time.perf_counter()
        tr = plan_oracle(cfg, model, obs, t0=[s == 0] * envs_per_step, prev_mean=prev, noise=noise)
        dt = time.perf_counter() - t
        prev = tr.mean
        if s >= warmup:
            times.append(dt)
    t_step = sum(times) / len(times)
    value = envs_per_step * cfg.num_samples * cfg.horizon / t_step
    return value, t_step, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload(WORKLOAD)
    envs = 2
    value, t_step, cores = cpu_reference_run(cfg, envs, args.steps, min(args.warmup, 3), budget_s=150.0)
    sample = f"{envs} of the {cfg.num_envs} environments per step, planned sequentially (the reference has no env axis)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: dog-run 5M model, num_samples=512, horizon=3, iterations=6 (reference algorithm, "
                               "host CPU, eager PyTorch fp32)", "envs_per_step": envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default=None, choices=["tcgen05x2pf", "tcgen05pp", "tcgen05x2", "tcgen05", "simt"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--envs", type=int, default=None, help="environments per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from tdmpc2_b200.tdmpc2 import TDMPC2
    from tdmpc2_b200.sharded import ShardedActor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the planner has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"

    over = {} if args.envs is None else {"num_envs": args.envs}
    cfg = workload(args.workload, iterations_effective=True, **over)
    E_local = cfg.num_envs                       # weak scaling: per-GPU work is fixed
    E_total = E_local * world
    agent = TDMPC2(cfg, device=dev, engine=args.engine)
    agent.load(synth_state_dict(cfg, seed=1))
    gen = torch.Generator(device=dev).manual_seed(3 + rank)
    agent.generator = gen
    obs_dim, A = cfg.obs_shape["state"][0], cfg.action_dim
    g = torch.Generator().manual_seed(2)
    obs_host_all = torch.randn(E_total, obs_dim, generator=g)
    obs_host = obs_host_all[rank * E_local:(rank + 1) * E_local].clone().pin_memory()
    obs_dev = obs_host.to(dev)
    task_dev = None
    if cfg.multitask:
        task_dev = (torch.arange(rank * E_local, (rank + 1) * E_local) % len(cfg.tasks)).to(torch.int32).to(dev)
    actions_host = torch.empty(E_total, A).pin_memory()
    gather_buf = torch.empty(E_total, A, device=dev)

    def step_device(t0):
        """plan() with inputs resident in HBM (+ the action all-gather when sharded)."""
        a = agent._plan(obs_dev, t0=t0, eval_mode=False, task=task_dev)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, a.contiguous())
            return gather_buf
        return a

    def step_e2e(t0):
        """The user-facing call: HOST observations in, HOST actions out."""
        o = obs_host.to(dev, non_blocking=True)
        a = agent._plan(o, t0=t0, eval_mode=False, task=task_dev)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, a.contiguous())
            a = gather_buf
        actions_host[: a.shape[0]].copy_(a, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return actions_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between CUDA events, barrier + synchronize on both sides, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn(False)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- warm-up (first call t0=True, then steady-state warm starts)
    step_device(True)
    for _ in range(args.warmup - 1):
        step_device(False)
    barrier()
    launches0 = agent.planner.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step_device, args.steps)
    launches = agent.planner.launches - launches0
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- dominant kernel: one CEM-iteration launch, timed alone with events on its stream
    from tdmpc2_b200.planner import draw_noise
    pl = agent.planner
    noise = draw_noise(cfg, E_local, dev, generator=gen)
    t0v = torch.zeros(E_local, dtype=torch.uint8, device=dev)
    prev = agent._prev_mean.reshape(E_local, cfg.horizon, A).contiguous()
    pl.prologue(obs_dev, task_dev, t0v, prev, noise.prior)
    its = [(noise.r[:, i].contiguous(), noise.pi[:, i].contiguous(), noise.qidx[:, i].contiguous())
           for i in range(cfg.iterations)]
    for a_ in its[:2]:
        pl.iterate(*a_)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        for a_ in its:
            pl.iterate(*a_)
    e1.record()
    torch.cuda.synchronize()
    ms_iter = e0.elapsed_time(e1) / (reps * cfg.iterations)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    traffic = None
    try:   # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture of this workload
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            if args.workload == "c2" and E_local == 256:
                traffic = json.load(f)["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    L, M, A_, T, B = cfg.latent_dim, cfg.mlp_dim, cfg.action_dim, cfg.task_dim, cfg.num_bins
    D = L + T + A_
    w = lambda i, h, o: i * h + h * h + h * o
    flops_iter = 2.0 * E_local * cfg.num_samples * (cfg.horizon * (w(D, M, B) + w(D, M, L)) + w(L + T, M, 2 * A_) + 2 * w(D, M, B))
    achieved = flops_iter / (ms_iter * 1e-3) / 1e12
    peak = float(peaks["bf16_tflops"])
    steps_per_plan = E_total * cfg.num_samples * cfg.horizon
    value = steps_per_plan / (ms_step * 1e-3)
    e2e_value = steps_per_plan / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: dog-run 5M model, {E_local} envs/GPU, num_samples={cfg.num_samples}, "
                               f"horizon={cfg.horizon}, iterations={cfg.iterations}" if args.workload == "c2"
                   else f"{args.workload}: {E_local} envs/GPU",
                   "global_envs": E_total, "parallelism": f"env-shard x{world}", "engine": agent.planner.engine_name,
                   "arithmetic": "3-pass fp16-split operands on tcgen05 kind::f16, fp32 accumulate (fp32-parity mode)",
                   "l2": "no flush: per-step inputs exceed L2 (fresh noise tensors, "
                         f"{4 * E_local * cfg.iterations * (cfg.horizon * (cfg.num_samples - cfg.num_pi_trajs) + cfg.num_samples) * A_ / 1e6:.0f} MB/step/GPU)",
                   "tflops_algorithmic": flops_per_env(cfg) * E_total / (ms_step * 1e-3) / 1e12},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(E_local * obs_dim * 4), "d2h_bytes_per_step": int(E_total * A_ * 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "plan_kernel<tcgen05> MODE_ITER (one CEM iteration)",
                     "ms_per_launch": ms_iter, "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({peak_src}, burst)",
                     "note": "achieved counts ALGORITHMIC flops (2 Q heads, 1x); the fp32-parity path issues 3 fp16 MMAs "
                             "per product, so its ceiling is peak/3"},
    }
    if world == 1 and not args.no_cpu_baseline:
        envs, n_steps = 2, 16
        v, t, cores = cpu_reference_run(cfg, envs, n_steps, 1, budget_s=20.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{envs} environments x up to {n_steps} plan() calls (20 s budget) of the same workload, sequential "
                                          f"(reference has no env axis), eager PyTorch fp32, {cores} threads"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
