"""The reference's evaluation loop over an environments axis (the caller of the planner).

`evaluate.py:71-83` of the reference runs `eval_episodes` episodes per task one after another, one `agent.act(obs,
t0=t==0, task=task_idx)` per environment step.  Here E environment instances advance in lock-step and every step is ONE
batched `agent.act(obs[E], t0[E], task[E])` -- one replay of the planner's CUDA graph for all of them.  Episodes are
handed to the E slots from a queue in the reference's order (task-major, then episode index); a slot whose episode ends
takes the next one and flags `t0` for itself only, so slots run out of phase and, on multi-task models, on different
tasks at the same time.  The statistics are the reference's: mean episode reward and mean `info['success']` per task,
and for multi-task models the normalised score of `evaluate.py:91,96` (success * 100 for `mw-*` tasks, reward / 10
otherwise, averaged over tasks).

Several GPUs: one process per GPU, each with its own agent and its own E environment instances.  Episodes are independent
units, so rank g of G takes every G-th episode of the queue and runs the same loop with NO collective in it; the
per-episode records are exchanged once at the end (`all_gather_object`, a few floats per episode) and every rank returns
the same merged statistics.

Environments are the reference's wrappers as `envs.make_env` returns them (`reset(task_idx=...)` -> obs tensor,
`step(action)` -> (obs, reward, done, info), `envs/wrappers/tensor.py`, `multitask.py`); nothing here imports them.
Video capture and the hydra entry point are not part of the planning path and are not provided.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass
class TaskResult:
    task: str
    episode_rewards: List[float] = field(default_factory=list)
    episode_successes: List[float] = field(default_factory=list)
    episode_lengths: List[int] = field(default_factory=list)

    @property
    def reward(self) -> float:                       # evaluate.py:88
        return float(sum(self.episode_rewards) / max(len(self.episode_rewards), 1))

    @property
    def success(self) -> float:                      # evaluate.py:89
        return float(sum(self.episode_successes) / max(len(self.episode_successes), 1))

    @property
    def score(self) -> float:                        # evaluate.py:91
        return self.success * 100 if self.task.startswith("mw-") else self.reward / 10


@dataclass
class _Slot:
    task_idx: Optional[int] = None
    episode: int = -1
    t: int = 0
    ep_reward: float = 0.0
    active: bool = False


@torch.no_grad()
def evaluate(agent, envs: Sequence[Any], eval_episodes: int, eval_mode: bool = False, verbose: bool = False,
             group: Optional["dist.ProcessGroup"] = None) -> Dict[str, Any]:
    """Evaluate `agent` (a `tdmpc2_b200.TDMPC2` built with `cfg.num_envs == len(envs)`) for `eval_episodes` episodes per
    task.  `eval_mode` is passed to `act()`; the reference's script leaves it at its default (False: the planner adds its
    final Gaussian noise, tdmpc2.py:203).  Under torch.distributed every rank passes its own agent and environments and
    gets the merged result.  Returns {"tasks": {name: TaskResult}, "normalized_score": float | None, "env_steps": int,
    "act_calls": int} (env_steps summed over ranks, act_calls the maximum over ranks)."""
    cfg = agent.cfg
    E = len(envs)
    if E < 1 or eval_episodes < 1:                                            # evaluate.py:42
        raise ValueError("Must evaluate at least 1 episode on at least 1 environment.")
    if E != int(getattr(agent, "num_envs", E)):
        raise ValueError(f"{E} environments for an agent built with cfg.num_envs={agent.num_envs}")
    multitask = bool(cfg.multitask)
    names = list(cfg.tasks) if multitask else [cfg.task]                      # evaluate.py:70
    results = {name: TaskResult(name) for name in names}
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    items = [(ti if multitask else None, ep) for ti in range(len(names)) for ep in range(eval_episodes)]
    queue = deque(items[rank::world])                                         # episodes are independent: no exchange in the loop
    records: List[tuple] = []                                                 # (task position, episode, reward, success, length)
    slots = [_Slot() for _ in range(E)]
    obs: List[Optional[torch.Tensor]] = [None] * E

    def start(i: int) -> None:
        s = slots[i]
        if not queue:
            s.active, obs[i] = False, None                                        # idle: planned on a zero observation, output ignored
            return
        s.task_idx, s.episode = queue.popleft()
        s.t, s.ep_reward, s.active = 0, 0.0, True
        obs[i] = torch.as_tensor(envs[i].reset(task_idx=s.task_idx), dtype=torch.float32)      # evaluate.py:76

    for i in range(E):
        start(i)
    filler = next((o for o in obs if o is not None), None)
    env_steps = act_calls = 0
    while any(s.active for s in slots):
        batch = torch.stack([o if o is not None else torch.zeros_like(filler) for o in obs])
        t0 = torch.tensor([s.t == 0 or not s.active for s in slots])          # idle slots plan from scratch and are ignored
        task = torch.tensor([s.task_idx or 0 for s in slots]) if multitask else None
        actions = agent.act(batch, t0=t0, eval_mode=eval_mode, task=task)     # evaluate.py:80, all environments at once
        act_calls += 1
        actions = actions.reshape(E, -1)
        for i, s in enumerate(slots):
            if not s.active:
                continue
            o, reward, done, info = envs[i].step(actions[i])                  # evaluate.py:81
            obs[i] = torch.as_tensor(o, dtype=torch.float32)
            s.ep_reward += float(reward)
            s.t += 1
            env_steps += 1
            if bool(done):
                records.append((s.task_idx or 0, s.episode, s.ep_reward, float(info.get("success", 0.0)), s.t))
                start(i)
    if world > 1:                                                             # the one exchange: per-episode records
        parts: List[Any] = [None] * world
        dist.all_gather_object(parts, (records, env_steps, act_calls), group=group)
        records = [r for p in parts for r in p[0]]
        env_steps, act_calls = sum(p[1] for p in parts), max(p[2] for p in parts)
    for ti, ep, rew, suc, length in sorted(records, key=lambda r: (r[0], r[1])):
        r = results[names[ti]]
        r.episode_rewards.append(rew)                                         # evaluate.py:86-87
        r.episode_successes.append(suc)
        r.episode_lengths.append(length)
    verbose = verbose and rank == 0
    for name in names:
        if verbose:
            print(f"  {name:<22}\tR: {results[name].reward:.01f}  \tS: {results[name].success:.02f}")
    score = float(sum(results[n].score for n in names) / len(names)) if multitask else None      # evaluate.py:96
    if verbose and multitask:
        print(f"Normalized score: {score:.02f}")
    return {"tasks": results, "normalized_score": score, "env_steps": env_steps, "act_calls": act_calls}
