"""Shared helpers for parity tests (tolerances are written where they are used)."""
import os

import numpy as np
import torch

from tdmpc2_b200.config import workload
from tdmpc2_b200.synth import synth_state_dict, state_dict_checksum

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    f = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    over = eval(str(f["overrides"]))  # repr() of a plain dict written by oracle/make_golden.py
    cfg = workload(str(f["workload"]), **over)
    sd = synth_state_dict(cfg, seed=int(f["weight_seed"]), perturb=bool(f["perturb"]),
                          emb_scale=float(f["emb_scale"]))
    if "term_bias" in f.files:      # episodic fixtures: the calibrated termination bias they were minted with
        sd["_termination.2.bias"] = torch.full_like(sd["_termination.2.bias"], float(f["term_bias"]))
    chk = state_dict_checksum(sd)
    assert abs(chk - float(f["weight_checksum"])) <= 1e-9 * abs(chk), \
        "synthetic weights differ from the ones the golden vectors were minted with (torch RNG drift?)"
    calls = []
    for i in range(int(f["n_calls"])):
        g = lambda k: f[f"c{i}_{k}"]
        task = int(g("task"))
        calls.append(dict(obs=torch.from_numpy(g("obs")), t0=bool(g("t0")), eval_mode=bool(g("eval_mode")),
                          task=None if task < 0 else task, seed=int(g("seed")),
                          prev_mean=torch.from_numpy(g("prev_mean")), action=torch.from_numpy(g("action")),
                          mean=torch.from_numpy(g("mean")), values=torch.from_numpy(g("values")),
                          elite_idx=torch.from_numpy(g("elite_idx"))))
    return cfg, sd, calls


def stable_positions(values, k, tol):
    """bool [..., k]: sorted-top-k positions whose value is further than `tol` from
    both neighbours in the sorted order (incl. the (k+1)-th value).  Only there is
    a bit-exact sorted top-k index well defined under fp32 re-association noise."""
    top = torch.topk(values, k + 1, dim=-1).values
    gaps = top[..., :-1] - top[..., 1:]                       # [..., k]; gaps[j] = v_j - v_{j+1}
    ok_next = gaps > tol
    ok_prev = torch.cat([torch.ones_like(ok_next[..., :1]), ok_next[..., :-1]], dim=-1)
    return ok_next & ok_prev


def boundary_separated(values, k, tol):
    """bool [...]: the k-th and (k+1)-th largest values differ by more than `tol`,
    i.e. the elite SET (hence the refit mean/std) is well defined."""
    top = torch.topk(values, k + 1, dim=-1).values
    return (top[..., k - 1] - top[..., k]) > tol
