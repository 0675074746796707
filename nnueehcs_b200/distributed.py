"""K-axis sharding of the UQ forward over the GPUs of one node.

The reference is single-GPU (no collective anywhere).  The B200-native
path shards the member axis -- ensemble members, MC-dropout passes or Delta-UQ anchors -- across
ranks (one process per GPU, ``torch.distributed`` over NCCL/NVLink): every rank sees all N samples
(x is ~20 B/sample), reduces its own members inside the fused kernel to per-sample
``(count, mean, M2)``, and the shards are combined with ONE collective -- an all-gather of the
``[2, N, out]`` float32 (mean, M2) slab -- followed by a Chan merge kernel
(``uq_moments_merge``).  Moments cross the wire rather than raw power sums because
``sum(y^2) - sum(y)^2/n`` cancels catastrophically in fp32 when std << |mean|, which is the
in-distribution case.  Philox masks are keyed by the *global* pass id, so the result does not
depend on the number of ranks.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def split_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split of ``range(total)``: returns (begin, count) of ``rank``."""
    begin = (total * rank) // world
    end = (total * (rank + 1)) // world
    return begin, end - begin


class KShard:
    def __init__(self, group: Optional["dist.ProcessGroup"] = None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("KShard needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def split(self, total: int, rank: Optional[int] = None) -> Tuple[int, int]:
        return split_range(total, self.world, self.rank if rank is None else rank)

    def counts(self, total: int) -> List[int]:
        return [self.split(total, r)[1] for r in range(self.world)]

    def exchange(self, mean: torch.Tensor, m2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """The one collective: all-gather every rank's (mean, M2) slab -> two ``[world, ...]``
        tensors ordered by rank."""
        slab = torch.stack([mean, m2]).contiguous()
        flat = torch.empty(self.world * slab.numel(), dtype=slab.dtype, device=slab.device)
        dist.all_gather_into_tensor(flat, slab.reshape(-1), group=self.group)
        gathered = flat.view((self.world,) + tuple(slab.shape))
        return gathered[:, 0], gathered[:, 1]

    def forward_owned(self, packed: "ops.PackedModel", x: torch.Tensor, mode: str, *,
                      local_members: int, precision: str = "fp32",
                      **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        """Each rank's ``packed`` holds ONLY its own shard of the members (an ensemble whose
        weights are distributed over the GPUs): reduce locally, exchange moments, merge."""
        mean, m2 = packed.forward(x, mode, total_members=local_members, precision=precision,
                                  output="moments", **kw)
        means, m2s = self.exchange(mean, m2)
        return ops.moments_merge(means.contiguous(), m2s.contiguous(),
                                 self.owned_counts(local_members, x.device))

    def owned_counts(self, local_members: int, device) -> List[float]:
        """Member count of every rank's shard (gathered once, then cached: it is static)."""
        cache = self.__dict__.setdefault("_owned_counts", {})
        if local_members not in cache:
            cnt = torch.tensor([float(local_members)], dtype=torch.float64, device=device)
            all_cnt = torch.empty(self.world, dtype=torch.float64, device=device)
            dist.all_gather_into_tensor(all_cnt, cnt, group=self.group)
            cache[local_members] = all_cnt.tolist()
        return cache[local_members]

    def forward(self, packed: "ops.PackedModel", x: torch.Tensor, mode: str, *, total_members: int,
                precision: str = "fp32", **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        begin, count = self.split(total_members)
        if count > 0:
            mean, m2 = packed.forward(x, mode, total_members=total_members, precision=precision,
                                      member_begin=begin, member_count=count, output="moments",
                                      **kw)
        else:  # more ranks than members: this rank contributes an empty shard
            mean = torch.zeros((x.shape[0], packed.d_out), dtype=torch.float32, device=x.device)
            m2 = torch.zeros_like(mean)
        means, m2s = self.exchange(mean, m2)
        counts = self.counts(total_members)
        live = [r for r, c in enumerate(counts) if c > 0]
        if len(live) != self.world:
            idx = torch.tensor(live, device=means.device)
            means, m2s = means.index_select(0, idx), m2s.index_select(0, idx)
            counts = [counts[r] for r in live]
        return ops.moments_merge(means.contiguous(), m2s.contiguous(), counts)
