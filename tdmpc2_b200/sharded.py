"""Environment-axis sharding over GPUs (one process per GPU, torch.distributed).

Environments are independent units: rank g plans the
contiguous block [g*E/G, (g+1)*E/G) with replicated weights and NO communication
inside the CEM loop; the only collective is one all-gather of the selected
actions [E/G, A] per plan() -- issued only when the environment axis is sharded.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs: int, rank: int, world: int) -> Tuple[int, int]:
    if num_envs % world != 0:
        raise ValueError(f"num_envs={num_envs} must be divisible by the number of ranks ({world})")
    per = num_envs // world
    return rank * per, (rank + 1) * per


class ShardedActor:
    """Wraps a per-rank `plan_local(obs_local, t0_local, task_local) -> actions_local`
    callable; `act()` takes / returns GLOBAL batches."""

    def __init__(self, plan_local: Callable, num_envs: int, group: Optional[dist.ProcessGroup] = None):
        self.plan_local = plan_local
        self.num_envs = num_envs
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_range(num_envs, self.rank, self.world)

    def local(self, x):
        if x is None or not torch.is_tensor(x) or x.ndim == 0:
            return x
        return x[self.lo:self.hi]

    def gather(self, a_local: torch.Tensor) -> torch.Tensor:
        """[E/G, A] of every rank -> [E, A] in rank order: the ONE collective of a sharded plan()."""
        if self.world == 1:
            return a_local
        out = torch.empty(self.num_envs, a_local.shape[-1], dtype=a_local.dtype, device=a_local.device)
        dist.all_gather_into_tensor(out, a_local.contiguous(), group=self.group)
        return out

    def act_local(self, obs_local, t0=False, task=None, **kw) -> torch.Tensor:
        """This rank's shard in (already sliced), GLOBAL actions out."""
        return self.gather(self.plan_local(obs_local, t0, task, **kw))

    def act(self, obs, t0=False, task=None, **kw) -> torch.Tensor:
        """GLOBAL batch in, GLOBAL actions out."""
        return self.act_local(self.local(obs), self.local(t0) if torch.is_tensor(t0) else t0, self.local(task), **kw)
