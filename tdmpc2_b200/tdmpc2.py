"""TDMPC2 agent with the reference's inference surface, planning on the B200 kernels.

Drop-in for the inference half of the reference class `TDMPC2`
(tdmpc2/tdmpc2.py:10-206): `TDMPC2(cfg)`, `.model`, `.cfg`, `.device`,
`._prev_mean`, `.discount`, `.load()`, `.save()`, `.act()`, `.plan`, `._plan()`,
`._estimate_value()` keep their names, argument meaning and return shapes, so
`evaluate.py:57-80` of the reference runs unchanged on it.
Training (`update`, optimisers, RunningScale) is out of scope (SURVEY.md 2.1 #1b).

New: an environments axis.  `obs [E, obs_dim]`, `t0 [E]`, `task [E]` plan E
independent environments in one call; E == 1 (1-D obs) is the reference API.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Union

import torch

from .config import Config, get_discount
from .planner import Noise, Planner, draw_noise
from .world_model import WorldModel, convert_legacy_checkpoint


class TDMPC2(torch.nn.Module):
    def __init__(self, cfg: Config, device: Union[str, torch.device, None] = None, engine: Optional[str] = None):
        super().__init__()
        self.cfg = cfg
        self.device = torch.device("cuda:0" if device is None else device)      # tdmpc2.py:20
        self.model = WorldModel(cfg).to(self.device)
        self.model.eval()                                                        # tdmpc2.py:32
        if not cfg.get("iterations_effective", False):
            self.cfg.iterations += 2 * int(cfg.action_dim >= 20)                # tdmpc2.py:34
            self.cfg.iterations_effective = True
        self.discount = torch.tensor(
            [get_discount(cfg, ep) for ep in cfg.episode_lengths], device=self.device
        ) if cfg.multitask else get_discount(cfg, cfg.episode_length)           # tdmpc2.py:35-37
        self.num_envs = int(cfg.get("num_envs", 1) or 1)
        shape = (cfg.horizon, cfg.action_dim) if self.num_envs == 1 else (self.num_envs, cfg.horizon, cfg.action_dim)
        self._prev_mean = torch.nn.Buffer(torch.zeros(*shape, device=self.device))   # tdmpc2.py:40
        self._engine = engine
        self._planner: Optional[Planner] = None
        self._weights_dirty = True
        self.generator: Optional[torch.Generator] = None     # None -> torch's default CUDA generator, like the reference
        self._use_graph = bool(cfg.get("cuda_graph", True))   # replay the launch chain as one CUDA graph (cfg.compile's role)
        # one environment: interleave the reference-order noise draws with the launches instead (TDMPC2_B200_E1_GRAPH=1: A/B knob)
        self._e1_interleaved = bool(cfg.get("e1_interleaved", True)) and os.environ.get("TDMPC2_B200_E1_GRAPH", "0") in ("", "0")

    # ------------------------------------------------------------------ planner plumbing
    @property
    def planner(self) -> Planner:
        if self._planner is None:
            self._planner = Planner(self.cfg, self.num_envs, self.device, engine=self._engine)
            self._weights_dirty = True
        if self._weights_dirty:
            self._planner.pack(self.model.state_dict())
            self._weights_dirty = False
        return self._planner

    def sync_weights(self) -> None:
        """Call after modifying `self.model`'s parameters in place."""
        self._weights_dirty = True

    @property
    def plan(self):
        # The reference wraps _plan in torch.compile(mode="reduce-overhead") (tdmpc2.py:45-55);
        # here the fused kernels are the compiled artefact, so `plan` is `_plan`.
        return self._plan

    def save(self, fp):
        torch.save({"model": self.model.state_dict()}, fp)                       # tdmpc2.py:72-79

    def load(self, fp):
        """tdmpc2.py:81-95: path or dict, optional {"model": ...} wrapper, legacy-key conversion."""
        if isinstance(fp, dict):
            state_dict = fp
        else:
            state_dict = torch.load(fp, map_location=self.device, weights_only=False)
        state_dict = state_dict["model"] if "model" in state_dict else state_dict
        state_dict = convert_legacy_checkpoint(self.model.state_dict(), dict(state_dict))
        self.model.load_state_dict(state_dict)
        self._weights_dirty = True

    # ------------------------------------------------------------------ inference API
    @torch.no_grad()
    def act(self, obs, t0=False, eval_mode=False, task=None):
        """tdmpc2.py:97-120.  obs [obs_dim] (CPU) -> action [A] (CPU); or batched
        obs [E, obs_dim], t0 bool | [E], task int | [E] -> [E, A]."""
        obs = obs.to(self.device, non_blocking=True)
        batched = obs.ndim == (4 if self.cfg.get("obs", "state") == "rgb" else 2)   # rgb: [C, 64, 64] per environment
        if not batched:
            obs = obs.unsqueeze(0)
        if task is not None and not torch.is_tensor(task):
            task = torch.tensor([task] if isinstance(task, int) else list(task), device=self.device)
        if self.cfg.mpc:
            out = self._plan(obs, t0=t0, eval_mode=eval_mode, task=task)
            return out.cpu()
        out = self._policy_action(obs, eval_mode=eval_mode, task=task)           # tdmpc2.py:116-120
        return (out if batched else out[0]).cpu()

    @torch.no_grad()
    def _policy_action(self, obs, eval_mode=False, task=None, eps: Optional[torch.Tensor] = None):
        """The non-MPC branch of act() (tdmpc2.py:116-120): a = pi(encode(obs)), or its mean in eval_mode
        (`info["mean"]` = tanh(mean), world_model.py:173).  Runs the encode + policy-prior kernel modes: trajectory 0,
        step 0 of the prior rollout IS pi(encode(obs)) with noise eps (zero noise gives the mean)."""
        cfg, E, dev = self.cfg, self.num_envs, self.device
        if cfg.num_pi_trajs < 1:
            raise NotImplementedError("the kernel policy path needs cfg.num_pi_trajs >= 1")
        rgb = cfg.get("obs", "state") == "rgb"
        obs = obs.to(dev, torch.float32)
        obs = (obs.reshape(E, *cfg.obs_shape["rgb"]) if rgb else obs.reshape(E, -1)).contiguous()
        taskv = None
        if cfg.multitask:
            if task is None:
                raise ValueError("multi-task model needs `task`")
            taskv = torch.as_tensor(task, device=dev).reshape(-1).to(torch.int32)
            taskv = taskv.expand(E).contiguous() if taskv.numel() == 1 else taskv.contiguous()
        pl = self.planner
        shift = None
        if rgb:                                                                     # ShiftAug's draw comes first (layers.py:55)
            shift = torch.randint(0, 7, (E, 2), device=dev, dtype=torch.float32, generator=self.generator)
        noise = torch.zeros(E, cfg.horizon, cfg.num_pi_trajs, cfg.action_dim, device=dev)
        if not eval_mode:
            noise[:, 0, 0] = torch.randn(E, cfg.action_dim, device=dev, generator=self.generator) if eps is None else eps.to(dev)
        ones, zeros = torch.ones(E, dtype=torch.uint8, device=dev), torch.zeros(E, cfg.horizon, cfg.action_dim, device=dev)
        if rgb:
            pl.prologue_latent(pl.encode_pixels(obs, shift), taskv, ones, zeros, noise)
        else:
            pl.prologue(obs, taskv, ones, zeros, noise)
        return pl.get_state()["pi_actions"][:, 0, 0].clone()

    @torch.no_grad()
    def _plan(self, obs, t0=False, eval_mode=False, task=None, noise: Optional[Noise] = None, return_trace=False):
        """tdmpc2.py:138-206 on the fused kernels.  obs [1|E, obs_dim] on self.device.
        Returns action [A] when planning a single environment (reference shape),
        else [E, A]."""
        cfg, E, dev = self.cfg, self.num_envs, self.device
        obs = obs.to(dev, torch.float32)
        if obs.ndim == (3 if cfg.get("obs", "state") == "rgb" else 1):
            obs = obs.unsqueeze(0)
        if obs.shape[0] != E:
            raise ValueError(f"obs has {obs.shape[0]} environments, agent was built for cfg.num_envs={E}")
        obs = obs.contiguous()
        if torch.is_tensor(t0):
            t0v = t0.to(dev).reshape(-1).to(torch.uint8)
            t0v = t0v.expand(E).contiguous() if t0v.numel() == 1 else t0v.contiguous()
        else:
            t0v = torch.full((E,), int(bool(t0)), dtype=torch.uint8, device=dev) if not isinstance(t0, (list, tuple)) \
                else torch.tensor([int(bool(x)) for x in t0], dtype=torch.uint8, device=dev)
        taskv = None
        if cfg.multitask:
            if task is None:
                raise ValueError("multi-task model needs `task`")
            taskv = torch.as_tensor(task, device=dev).reshape(-1).to(torch.int32)
            taskv = taskv.expand(E).contiguous() if taskv.numel() == 1 else taskv.contiguous()
        prev = self._prev_mean.reshape(E, cfg.horizon, cfg.action_dim).contiguous()
        trace = None
        if noise is None and not return_trace and self._use_graph:
            # steady state: replay the captured prologue -> I x iter -> epilogue chain (tdmpc2.py:45-55 replays a
            # `reduce-overhead` graph); noise is drawn into the graph's static buffers
            try:
                if E == 1 and self._e1_interleaved and not self.planner.philox:
                    # one environment draws in the reference's order (29 small launches for I = 8): issue each draw right
                    # before its consumer instead of all of them ahead of a graph replay (planner.plan_interleaved)
                    action, new_mean = self.planner.plan_interleaved(obs, taskv, t0v, prev, eval_mode=eval_mode,
                                                                     generator=self.generator)
                else:
                    action, new_mean = self.planner.plan_graphed(obs, taskv, t0v, prev, eval_mode=eval_mode,
                                                                 generator=self.generator)
            except RuntimeError as e:                      # capture unsupported here: keep the eager launch chain
                import warnings
                warnings.warn(f"CUDA-graph capture of the plan chain failed ({e}); launching eagerly")
                self._use_graph = False
                return self._plan(obs, t0=t0, eval_mode=eval_mode, task=task)
        else:
            if noise is None:
                noise = draw_noise(cfg, E, dev, eval_mode=eval_mode, generator=self.generator)
            elif eval_mode:
                noise = Noise(noise.prior, noise.r, noise.pi, noise.qidx, noise.expo, None, noise.shift)
            action, new_mean, trace = self.planner.plan(obs, taskv, t0v, prev, noise, trace=return_trace)
        self._prev_mean.copy_(new_mean.reshape(self._prev_mean.shape))           # tdmpc2.py:205
        out = action[0] if E == 1 else action
        return (out, trace) if return_trace else out

    @torch.no_grad()
    def _estimate_value(self, z, actions, task, eps_pi=None, qidx=None):
        """tdmpc2.py:122-136.  z [N, L], actions [H, N, A] -> [N, 1]  (E == 1), or
        z [E, N, L], actions [E, H, N, A] -> [E, N, 1]."""
        cfg, E, dev = self.cfg, self.num_envs, self.device
        single = z.ndim == 2
        zb = (z.unsqueeze(0) if single else z).to(dev, torch.float32).contiguous()
        ab = (actions.unsqueeze(0) if single else actions).to(dev, torch.float32).contiguous()
        if eps_pi is None:
            eps_pi = torch.randn(E, cfg.num_samples, cfg.action_dim, device=dev, generator=self.generator)
        if qidx is None:
            qidx = torch.rand(E, cfg.num_q, device=dev, generator=self.generator).argsort(-1)[:, :2]
        taskv = None
        if cfg.multitask:
            taskv = torch.as_tensor(task, device=dev).reshape(-1).to(torch.int32)
            taskv = taskv.expand(E).contiguous() if taskv.numel() == 1 else taskv.contiguous()
        v = self.planner.estimate_value(zb, ab, taskv, eps_pi.reshape(E, cfg.num_samples, cfg.action_dim).contiguous(),
                                        qidx.reshape(E, 2).to(torch.int32).contiguous())
        return v[0].unsqueeze(-1) if single else v.unsqueeze(-1)
