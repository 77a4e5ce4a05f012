"""WorldModel with the reference's attribute names and state_dict() key layout,
without the tensordict dependency.

Mirrors (does not copy) tdmpc2/common/world_model.py:12-216 and the building
blocks of tdmpc2/common/layers.py:8-33,74-164 of the reference.  It exists for
two reasons:

  * drop-in checkpoint compatibility: `state_dict()` / `load_state_dict()` use
    exactly the reference's keys, including the
    stacked `_Qs.params.*`, `_detach_Qs_params.*`, `_target_Qs_params.*` and
    the tensordict metadata entries `__batch_size` / `__device`;
  * the non-planning forward methods (`encode/next/reward/pi/Q`) a reference
    user may call directly.  These are plain PyTorch and are NOT the planning
    hot path: TDMPC2.plan() runs the fused sm_100a kernels through the C ABI.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .synth import QS_PREFIXES, head_layout

_LAYER_PARAM_NAMES = ("weight", "bias", "ln.weight", "ln.bias")


class SimNorm(nn.Module):
    """Simplicial normalisation: softmax over groups of `simnorm_dim` (layers.py:74-88)."""

    def __init__(self, cfg):
        super().__init__()
        self.dim = cfg.simnorm_dim

    def forward(self, x):
        shp = x.shape
        return F.softmax(x.view(*shp[:-1], -1, self.dim), dim=-1).view(*shp)

    def __repr__(self):
        return f"SimNorm(dim={self.dim})"


class NormedLinear(nn.Linear):
    """Linear -> (Dropout) -> LayerNorm -> activation (layers.py:94-111)."""

    def __init__(self, *args, dropout=0., act=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.ln = nn.LayerNorm(self.out_features)
        self.act = act if act is not None else nn.Mish(inplace=False)
        self.dropout = nn.Dropout(dropout, inplace=False) if dropout else None

    def forward(self, x):
        x = super().forward(x)
        if self.dropout:
            x = self.dropout(x)
        return self.act(self.ln(x))


def mlp(in_dim, mlp_dims, out_dim, act=None, dropout=0.):
    """layers.py:121-133."""
    if isinstance(mlp_dims, int):
        mlp_dims = [mlp_dims]
    dims = [in_dim] + list(mlp_dims) + [out_dim]
    seq = nn.ModuleList()
    for i in range(len(dims) - 2):
        seq.append(NormedLinear(dims[i], dims[i + 1], dropout=dropout * (i == 0)))
    seq.append(NormedLinear(dims[-2], dims[-1], act=act) if act else nn.Linear(dims[-2], dims[-1]))
    return nn.Sequential(*seq)


def weight_init(m):
    """common/init.py:4-11."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Embedding):
        nn.init.uniform_(m.weight, -0.02, 0.02)


class StackedParams(nn.Module):
    """Stand-in for tensordict's TensorDictParams: a flat set of tensors named
    '<layer>.<param>' with a leading [num_q] dimension, serialised under the
    same keys tensordict uses (plus its `__batch_size` / `__device` entries)."""

    def __init__(self, tensors: Dict[str, torch.Tensor], num_q: int, as_params: bool = True):
        super().__init__()
        self._num_q = num_q
        self._names = list(tensors.keys())
        for k, v in tensors.items():
            p = v if isinstance(v, nn.Parameter) else nn.Parameter(v, requires_grad=as_params)
            self.register_parameter(self._mangle(k), p)

    @staticmethod
    def _mangle(k: str) -> str:
        return "p_" + k.replace(".", "_")

    def __getitem__(self, k):
        if isinstance(k, tuple):
            k = ".".join(k)
        return getattr(self, self._mangle(k))

    def keys(self):
        return list(self._names)

    def items(self):
        return [(k, self[k]) for k in self._names]

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for k in self._names:
            p = self[k]
            destination[prefix + k] = p if keep_vars else p.detach()
        first = self[self._names[0]]
        destination[prefix + "__batch_size"] = torch.Size([self._num_q])
        destination[prefix + "__device"] = first.device

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        for k in self._names:
            key = prefix + k
            if key in state_dict:
                src = state_dict[key]
                dst = self[k]
                if tuple(src.shape) != tuple(dst.shape):
                    error_msgs.append(f"size mismatch for {key}: {tuple(src.shape)} vs {tuple(dst.shape)}")
                else:
                    with torch.no_grad():
                        dst.copy_(src)
            elif strict:
                missing_keys.append(key)
        known = {prefix + k for k in self._names} | {prefix + "__batch_size", prefix + "__device"}
        if strict:
            for key in state_dict:
                if key.startswith(prefix) and key not in known:
                    unexpected_keys.append(key)


class Ensemble(nn.Module):
    """Vectorised ensemble of `num_q` identical MLPs (layers.py:8-33).  Parameters
    are stacked along dim 0; forward is a batched matmul chain (no vmap)."""

    def __init__(self, modules):
        super().__init__()
        self._n = len(modules)
        self._repr = str(modules[0])
        self._has_ln = []
        tensors: Dict[str, torch.Tensor] = {}
        for li, layer in enumerate(modules[0]):
            has_ln = isinstance(layer, NormedLinear)
            self._has_ln.append(has_ln)
            names = _LAYER_PARAM_NAMES if has_ln else _LAYER_PARAM_NAMES[:2]
            for nm in names:
                get = lambda m: (m[li].ln.weight if nm == "ln.weight" else m[li].ln.bias if nm == "ln.bias"
                                 else getattr(m[li], nm))
                tensors[f"{li}.{nm}"] = torch.stack([get(m).detach() for m in modules])
        self.params = StackedParams(tensors, self._n)

    def __len__(self):
        return self._n

    def forward_with(self, params, x):
        """x [rows, in] -> [num_q, rows, out]; Dropout is inert in eval mode."""
        h = x.unsqueeze(0).expand(self._n, *x.shape)
        last = len(self._has_ln) - 1
        for li, has_ln in enumerate(self._has_ln):
            w, b = params[f"{li}.weight"], params[f"{li}.bias"]
            h = torch.baddbmm(b.unsqueeze(1), h, w.transpose(1, 2))
            if has_ln:
                g, beta = params[f"{li}.ln.weight"], params[f"{li}.ln.bias"]
                mu = h.mean(-1, keepdim=True)
                var = h.var(-1, unbiased=False, keepdim=True)
                h = (h - mu) / torch.sqrt(var + 1e-5) * g.unsqueeze(1) + beta.unsqueeze(1)
                h = F.mish(h) if li != last else h
        return h

    def forward(self, x):
        return self.forward_with(self.params, x)

    def __repr__(self):
        return f"Vectorized {len(self)}x " + self._repr


def two_hot_inv(x, cfg):
    """math.py:74-83."""
    if cfg.num_bins == 0:
        return x
    if cfg.num_bins == 1:
        return torch.sign(x) * (torch.exp(torch.abs(x)) - 1)
    bins = torch.linspace(cfg.vmin, cfg.vmax, cfg.num_bins, device=x.device, dtype=x.dtype)
    x = torch.sum(F.softmax(x, dim=-1) * bins, dim=-1, keepdim=True)
    return torch.sign(x) * (torch.exp(torch.abs(x)) - 1)


class WorldModel(nn.Module):
    """TD-MPC2 implicit world model (world_model.py:12-216), state-observation configs."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        if cfg.get("obs", "state") != "state":
            raise NotImplementedError("pixel encoder is out of scope for the planning path (SURVEY.md section 8(f))")
        if cfg.multitask:
            self._task_emb = nn.Embedding(len(cfg.tasks), cfg.task_dim, max_norm=1)
            self.register_buffer("_action_masks", torch.zeros(len(cfg.tasks), cfg.action_dim))
            for i in range(len(cfg.tasks)):
                self._action_masks[i, :cfg.action_dims[i]] = 1.
        lay = head_layout(cfg)
        enc_dims = lay["_encoder.state"]["dims"]
        self._encoder = nn.ModuleDict({"state": mlp(enc_dims[0][0], [d[1] for d in enc_dims[:-1]], cfg.latent_dim,
                                                    act=SimNorm(cfg))})
        D = cfg.latent_dim + cfg.action_dim + cfg.task_dim
        self._dynamics = mlp(D, 2 * [cfg.mlp_dim], cfg.latent_dim, act=SimNorm(cfg))
        self._reward = mlp(D, 2 * [cfg.mlp_dim], max(cfg.num_bins, 1))
        self._termination = mlp(cfg.latent_dim + cfg.task_dim, 2 * [cfg.mlp_dim], 1) if cfg.episodic else None
        self._pi = mlp(cfg.latent_dim + cfg.task_dim, 2 * [cfg.mlp_dim], 2 * cfg.action_dim)
        qs = [mlp(D, 2 * [cfg.mlp_dim], max(cfg.num_bins, 1), dropout=cfg.dropout).apply(weight_init)
              for _ in range(cfg.num_q)]
        self._Qs = Ensemble(qs)
        self.apply(weight_init)
        with torch.no_grad():                       # world_model.py:32
            self._reward[-1].weight.fill_(0)
            self._Qs.params["2.weight"].fill_(0)
        self.register_buffer("log_std_min", torch.tensor(float(cfg.log_std_min)))
        self.register_buffer("log_std_dif", torch.tensor(float(cfg.log_std_max)) - self.log_std_min)
        self.init()

    def init(self):
        """world_model.py:38-53: detached view (shared storage) and target copy of the Q params."""
        shared = {k: v for k, v in self._Qs.params.items()}
        self._detach_Qs_params = StackedParams(shared, self.cfg.num_q)
        self._target_Qs_params = StackedParams({k: v.detach().clone() for k, v in shared.items()},
                                               self.cfg.num_q, as_params=False)

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        self.init()
        return self

    @property
    def total_params(self):
        seen, tot = set(), 0
        for p in self.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p)); tot += p.numel()
        return tot

    def __repr__(self):
        return f"TD-MPC2 World Model (B200 planner build)\nLearnable parameters: {self.total_params:,}"

    def soft_update_target_Q(self):
        with torch.no_grad():
            for k, v in self._target_Qs_params.items():
                v.lerp_(self._detach_Qs_params[k], self.cfg.tau)

    # ---- forward methods (plain PyTorch; not the planning hot path) -------------------------
    def task_emb(self, x, task):
        if isinstance(task, int):
            task = torch.tensor([task], device=x.device)
        emb = self._task_emb(task.long())
        if x.ndim == 3:
            emb = emb.unsqueeze(0).repeat(x.shape[0], 1, 1)
        elif emb.shape[0] == 1:
            emb = emb.repeat(x.shape[0], 1)
        return torch.cat([x, emb], dim=-1)

    def encode(self, obs, task):
        if self.cfg.multitask:
            obs = self.task_emb(obs, task)
        return self._encoder[self.cfg.obs](obs)

    def next(self, z, a, task):
        if self.cfg.multitask:
            z = self.task_emb(z, task)
        return self._dynamics(torch.cat([z, a], dim=-1))

    def reward(self, z, a, task):
        if self.cfg.multitask:
            z = self.task_emb(z, task)
        return self._reward(torch.cat([z, a], dim=-1))

    def pi(self, z, task, eps: Optional[torch.Tensor] = None):
        """world_model.py:144-184.  Returns (action, info) with the reference's info keys."""
        if self.cfg.multitask:
            z = self.task_emb(z, task)
        mean, log_std = self._pi(z).chunk(2, dim=-1)
        log_std = self.log_std_min + 0.5 * self.log_std_dif * (torch.tanh(log_std) + 1)
        eps = torch.randn_like(mean) if eps is None else eps
        if self.cfg.multitask:
            m = self._action_masks[task]
            mean, log_std, eps = mean * m, log_std * m, eps * m
            action_dims = self._action_masks.sum(-1)[task].unsqueeze(-1)
        else:
            action_dims = None
        log_prob = (-0.5 * eps.pow(2) - log_std - 0.9189385175704956).sum(-1, keepdim=True)
        size = eps.shape[-1] if action_dims is None else action_dims
        scaled_log_prob = log_prob * size
        action = mean + eps * log_std.exp()
        mean, action = torch.tanh(mean), torch.tanh(action)
        log_prob = log_prob - torch.log(F.relu(1 - action.pow(2)) + 1e-6).sum(-1, keepdim=True)
        entropy_scale = scaled_log_prob / (log_prob + 1e-8)
        info = {"mean": mean, "log_std": log_std, "action_prob": 1., "entropy": -log_prob,
                "scaled_entropy": -log_prob * entropy_scale}
        return action, info

    def Q(self, z, a, task, return_type="min", target=False, detach=False, qidx=None):
        assert return_type in {"min", "avg", "all"}
        if self.cfg.multitask:
            z = self.task_emb(z, task)
        z = torch.cat([z, a], dim=-1)
        params = self._target_Qs_params if target else self._detach_Qs_params if detach else self._Qs.params
        out = self._Qs.forward_with(params, z)
        if return_type == "all":
            return out
        if qidx is None:
            qidx = torch.randperm(self.cfg.num_q, device=out.device)[:2]
        Q = two_hot_inv(out[qidx], self.cfg)
        return Q.min(0).values if return_type == "min" else Q.sum(0) / 2


def convert_legacy_checkpoint(target_state_dict, source_state_dict):
    """Accept checkpoints written by the reference's pre-torch.compile API
    (behaviour of layers.api_model_conversion, layers.py:167-221): there the Q
    ensemble was a ParameterList `_Qs.params.<n>` / `_target_Qs.params.<n>` with
    n = 4*layer + {0: weight, 1: bias, 2: ln.weight, 3: ln.bias}."""
    if "_detach_Qs_params.0.weight" in source_state_dict:
        return source_state_dict
    out = {}
    for key, val in source_state_dict.items():
        for old_prefix, new_prefixes in (("_Qs.params.", ("_Qs.params.", "_detach_Qs_params.")),
                                         ("_target_Qs.params.", ("_target_Qs_params.",))):
            if key.startswith(old_prefix):
                n = int(key[len(old_prefix):])
                name = f"{n // 4}.{_LAYER_PARAM_NAMES[n % 4]}"
                for npfx in new_prefixes:
                    out[npfx + name] = val
                break
        else:
            if "Qs" in key:
                raise AssertionError(f"key {key} contains 'Qs'")
            out[key] = val
    for pfx in QS_PREFIXES:
        for meta in ("__batch_size", "__device"):
            if pfx + meta in target_state_dict:
                out[pfx + meta] = target_state_dict[pfx + meta]
    for key in target_state_dict:
        if "Qs" in key and key not in out:
            raise AssertionError(f"key {key} not in converted checkpoint")
    for key in ("log_std_min", "log_std_dif", "_action_masks"):
        if key in target_state_dict:
            out[key] = target_state_dict[key]
    return out
