"""The vectorised evaluation loop (tdmpc2_b200/evaluate.py) against the reference's sequential loop (evaluate.py:71-96
restated below) on scripted environments and a scripted agent -- host logic only, no kernels."""
from collections import defaultdict
from types import SimpleNamespace

import pytest
import torch

from tdmpc2_b200.evaluate import evaluate


class ScriptedEnv:
    """Reference-wrapper call shape (envs/wrappers/tensor.py:36-46).  Episode length and rewards depend on the task, on a
    per-(task) episode counter shared by all instances (so that which slot runs an episode does not matter) and on the
    actions taken."""
    counters = defaultdict(int)

    def __init__(self, obs_dim=5, success_tasks=(1,)):
        self.obs_dim, self.success_tasks = obs_dim, success_tasks

    def reset(self, task_idx=None):
        self.task = task_idx
        key = -1 if task_idx is None else task_idx
        self.episode = ScriptedEnv.counters[key]
        ScriptedEnv.counters[key] += 1
        self.length = 3 + (self.episode * 2 + 3 * (key + 1)) % 5
        self.t = 0
        return self._obs()

    def _obs(self):
        o = torch.zeros(self.obs_dim)
        o[0], o[1], o[2], o[3] = float(-1 if self.task is None else self.task), float(self.episode), float(self.t), 1.0
        return o

    def step(self, action):
        assert isinstance(action, torch.Tensor) and action.device.type == "cpu" and action.ndim == 1
        self.t += 1
        reward = torch.tensor(float(action[0]) + 0.25 * self.t, dtype=torch.float32)
        info = defaultdict(float, success=float((self.task in self.success_tasks) and self.episode % 2 == 0))
        return self._obs(), reward, self.t >= self.length, info


class ScriptedAgent:
    """act() is a pure function of (obs, t0, task) so that batched and sequential runs must agree exactly; it checks
    that t0 is raised exactly at t == 0 and that the task routed to a row is the row's own."""

    def __init__(self, cfg, num_envs):
        self.cfg, self.num_envs, self.calls = cfg, num_envs, 0

    def _one(self, obs, t0, task):
        assert bool(t0) == (float(obs[2]) == 0.0), "t0 must be raised on the first step of an episode only"
        if self.cfg.multitask:
            assert int(task) == int(obs[0])
        a = torch.zeros(self.cfg.action_dim)
        a[0] = obs[0] * 0.5 + obs[1] * 0.125 + obs[2] * 0.0625 + (1.0 if bool(t0) else 0.0)
        return a

    def act(self, obs, t0=False, eval_mode=False, task=None):
        self.calls += 1
        if obs.ndim == 1:
            return self._one(obs, t0, task)
        assert obs.shape[0] == self.num_envs
        # idle slots arrive as zero rows (real observations carry o[3] == 1) with t0 raised; their output is discarded
        assert all(bool(t0[i]) for i in range(obs.shape[0]) if float(obs[i, 3]) == 0.0)
        return torch.stack([torch.full((self.cfg.action_dim,), float("nan")) if float(obs[i, 3]) == 0.0
                            else self._one(obs[i], t0[i], None if task is None else task[i]) for i in range(obs.shape[0])])


def sequential(agent, env, cfg, eval_episodes):
    """evaluate.py:70-96 without the video branch."""
    scores, out = [], {}
    tasks = cfg.tasks if cfg.multitask else [cfg.task]
    for task_idx, task in enumerate(tasks):
        if not cfg.multitask:
            task_idx = None
        ep_rewards, ep_successes = [], []
        for i in range(eval_episodes):
            obs, done, ep_reward, t = env.reset(task_idx=task_idx), False, 0, 0
            while not done:
                action = agent.act(obs, t0=t == 0, task=task_idx)
                obs, reward, done, info = env.step(action)
                ep_reward += reward
                t += 1
            ep_rewards.append(float(ep_reward))
            ep_successes.append(info["success"])
        out[task] = (sum(ep_rewards) / len(ep_rewards), sum(ep_successes) / len(ep_successes))
        if cfg.multitask:
            scores.append(out[task][1] * 100 if task.startswith("mw-") else out[task][0] / 10)
    return out, (sum(scores) / len(scores) if cfg.multitask else None)


@pytest.mark.parametrize("E", [1, 3, 8])
@pytest.mark.parametrize("multitask", [False, True])
def test_batched_loop_reproduces_the_sequential_loop(E, multitask):
    cfg = SimpleNamespace(multitask=multitask, task="walker-run", tasks=["walker-run", "mw-door-open", "cheetah-run"], action_dim=4)
    episodes = 5
    ScriptedEnv.counters.clear()
    want, want_score = sequential(ScriptedAgent(cfg, 1), ScriptedEnv(), cfg, episodes)
    ScriptedEnv.counters.clear()
    agent = ScriptedAgent(cfg, E)
    got = evaluate(agent, [ScriptedEnv() for _ in range(E)], episodes)
    names = cfg.tasks if multitask else [cfg.task]
    total_steps = 0
    for name in names:
        r = got["tasks"][name]
        assert len(r.episode_rewards) == episodes and len(r.episode_lengths) == episodes
        # episodes finish in a different order than they were started in: compare the statistics the reference prints
        assert r.reward == pytest.approx(want[name][0], rel=1e-6) and r.success == pytest.approx(want[name][1])
        total_steps += sum(r.episode_lengths)
    assert got["env_steps"] == total_steps and got["act_calls"] == agent.calls
    assert got["act_calls"] <= total_steps and (E == 1 or got["act_calls"] < total_steps)      # batching happened
    if multitask:
        assert got["normalized_score"] == pytest.approx(want_score, rel=1e-6)
        assert got["tasks"]["mw-door-open"].score == pytest.approx(got["tasks"]["mw-door-open"].success * 100)
    else:
        assert got["normalized_score"] is None


def test_argument_checks():
    cfg = SimpleNamespace(multitask=False, task="t", tasks=["t"], action_dim=2)
    with pytest.raises(ValueError):
        evaluate(ScriptedAgent(cfg, 2), [ScriptedEnv()], 1)                 # agent built for another batch size
    with pytest.raises(ValueError):
        evaluate(ScriptedAgent(cfg, 1), [ScriptedEnv()], 0)                 # evaluate.py:42


# ------------------------------------------------------------------------------------------------ two ranks over gloo
def _eval_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = SimpleNamespace(multitask=True, task="x", tasks=["walker-run", "mw-door-open", "cheetah-run"], action_dim=4)
        ScriptedEnv.counters.clear()
        agent = ScriptedAgent(cfg, 2)
        out = evaluate(agent, [ScriptedEnv(), ScriptedEnv()], 5)
        q.put((rank, {n: (r.episode_rewards, r.episode_successes, r.episode_lengths) for n, r in out["tasks"].items()},
               out["normalized_score"], out["env_steps"], out["act_calls"], agent.calls))
    finally:
        dist.destroy_process_group()


def test_two_ranks_split_the_episode_queue_and_merge_once():
    """world_size 2 over gloo: every rank runs its own two environments on every second episode of the queue, no
    collective in the loop, one all_gather_object at the end; both ranks return the same merged statistics, with
    eval_episodes episodes per task."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, t0_, s0, steps0, calls0, own0), (_, t1_, s1, steps1, calls1, own1) = res
    assert t0_ == t1_ and s0 == s1 and steps0 == steps1 and calls0 == calls1
    for name, (rew, suc, lens) in t0_.items():
        assert len(rew) == len(suc) == len(lens) == 5
    assert steps0 == sum(sum(v[2]) for v in t0_.values())
    assert calls0 == max(own0, own1) and own0 + own1 < steps0          # each rank batched its own two environments
