"""K-axis sharding of the UQ forward over the GPUs of one node.

The reference is single-GPU (no collective anywhere).  The B200-native
path shards the member axis -- ensemble members, MC-dropout passes or Delta-UQ anchors -- across
ranks (one process per GPU, ``torch.distributed`` over NCCL/NVLink): every rank sees all N samples
(x is ~20 B/sample), reduces its own members inside the fused kernel to per-sample
``(count, mean, M2)`` written straight into one ``[mean | M2]`` slab, and the shards are combined
reduce-scatter style: an all-to-all hands rank r the r-th slice of the rows of every shard, rank r
Chan-merges its slice (``uq_moments_merge_ex``) and an all-gather of the finished (mean, std)
slices rebuilds the result on every rank -- ``2 (G-1)/G`` slabs per GPU over NVLink instead of the
``G-1`` slabs (plus a G-slab merge read) of a plain all-gather of the moments.  Moments cross the
wire rather than raw power sums because
``sum(y^2) - sum(y)^2/n`` cancels catastrophically in fp32 when std << |mean|, which is the
in-distribution case.  Philox masks are keyed by the *global* pass id, so the result does not
depend on the number of ranks.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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
        """All-gather of every rank's (mean, M2) -> two ``[world, ...]`` tensors ordered by rank
        (the simple form of the exchange; ``combine`` is what the forwards use)."""
        slab = torch.stack([mean, m2]).contiguous()
        flat = torch.empty(self.world * slab.numel(), dtype=slab.dtype, device=slab.device)
        dist.all_gather_into_tensor(flat, slab.reshape(-1), group=self.group)
        gathered = flat.view((self.world,) + tuple(slab.shape))
        return gathered[:, 0], gathered[:, 1]

    # ---- reduce-scatter style combine ----------------------------------------------------------
    def slice_len(self, length: int) -> int:
        """Elements of one rank's slice of a ``length``-element result (the last may be padded)."""
        return -(-length // self.world)

    def new_slab(self, length: int, device) -> torch.Tensor:
        """``[2, world * slice_len]`` float32: row 0 takes the means, row 1 the M2 of this rank's
        members (first ``length`` entries; the padding is never read back)."""
        return torch.empty((2, self.world * self.slice_len(length)), dtype=torch.float32,
                           device=device)

    def combine(self, slab: torch.Tensor, counts: Sequence[float], merge=None) -> torch.Tensor:
        """``slab`` = this rank's ``[mean | M2]`` (``new_slab``) -> ``[2, world * slice_len]`` holding
        the merged (mean, unbiased std) of all shards, identical on every rank.
        all-to-all of the row slices -> Chan merge of slice ``rank`` -> all-gather of the slices."""
        world, cap = self.world, slab.shape[1] // self.world
        merge = ops.moments_merge_strided if merge is None else merge
        if world == 1:
            return merge(slab.view(1, 2, cap), list(counts))
        # [2][world][cap] -> [world][2][cap]: destination-major for the all-to-all
        send = slab.view(2, world, cap).transpose(0, 1).contiguous()
        recv = _all_to_all(send.view(-1), [2 * cap] * world, [2 * cap] * world, self.group)
        fin = merge(recv.view(world, 2, cap), list(counts))             # [2, cap]
        out = torch.empty((2, world * cap), dtype=slab.dtype, device=slab.device)
        for row in range(2):
            if slab.is_cuda:
                dist.all_gather_into_tensor(out[row], fin[row].contiguous(), group=self.group)
            else:  # gloo (CPU tests)
                dist.all_gather(list(out[row].view(world, cap).unbind(0)), fin[row].contiguous(),
                                group=self.group)
        return out

    def _forward_and_combine(self, packed, x, mode, counts, merge=None, **fkw):
        n, d = x.shape[0], packed.d_out
        length = n * d
        slab = self.new_slab(length, x.device)
        if fkw.pop("_skip", False):   # more ranks than members: an empty shard (count 0)
            slab.zero_()
        else:
            packed.forward_into(x, mode, slab[0, :length].view(n, d), slab[1, :length].view(n, d),
                                output="moments", **fkw)
        out = self.combine(slab, counts, merge)
        return out[0, :length].view(n, d), out[1, :length].view(n, d)

    def forward_owned(self, packed: "ops.PackedModel", x: torch.Tensor, mode: str, *,
                      local_members: int, precision: str = "fp32", merge=None,
                      **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        """Each rank's ``packed`` holds ONLY its own shard of the members (an ensemble whose
        weights are distributed over the GPUs): reduce locally, exchange moments, merge."""
        return self._forward_and_combine(packed, x, mode, self.owned_counts(local_members, x.device),
                                         merge, total_members=local_members, precision=precision,
                                         **kw)

    def owned_counts(self, local_members: int, device) -> List[float]:
        """Member count of every rank's shard (gathered once, then cached: it is static)."""
        cache = self.__dict__.setdefault("_owned_counts", {})
        if local_members not in cache:
            cnt = torch.tensor([float(local_members)], dtype=torch.float64, device=device)
            all_cnt = torch.empty(self.world, dtype=torch.float64, device=device)
            if cnt.is_cuda:
                dist.all_gather_into_tensor(all_cnt, cnt, group=self.group)
            else:
                dist.all_gather(list(all_cnt.view(self.world, 1).unbind(0)), cnt, group=self.group)
            cache[local_members] = all_cnt.tolist()
        return cache[local_members]

    def forward(self, packed: "ops.PackedModel", x: torch.Tensor, mode: str, *, total_members: int,
                precision: str = "fp32", merge=None, **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        begin, count = self.split(total_members)
        return self._forward_and_combine(packed, x, mode, self.counts(total_members), merge,
                                         total_members=total_members, precision=precision,
                                         member_begin=begin, member_count=count,
                                         _skip=(count == 0), **kw)


class NShard:
    """Sample-axis sharding of the fused forward (the alternative to the
    K axis, for K smaller than the number of GPUs).  Every rank holds the same ``x`` and the whole
    model, runs ALL members on its contiguous slice of the rows, and one all-gather of the
    ``[mean | second output]`` rows rebuilds the full result on every rank -- no moment algebra.
    Set ``model.uq_shard = NShard()`` on a mirror wrapper, exactly like ``KShard``.

    Native MC-dropout masks are a function of (pass, layer, GLOBAL row, feature): each rank passes
    the global index of its first row (``row_base``), so the bits -- and the result -- do not depend
    on the number of ranks, exactly as under ``KShard``.  Injected masks are laid out for the whole
    batch and are not supported here."""

    def __init__(self, group: Optional["dist.ProcessGroup"] = None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("NShard needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def rows(self, n: int, rank: Optional[int] = None) -> Tuple[int, int]:
        return split_range(n, self.world, self.rank if rank is None else rank)

    def gather_rows(self, local: torch.Tensor, n: int) -> torch.Tensor:
        """``local``: this rank's ``[count, ...]`` rows -> the full ``[n, ...]`` tensor (rank order)."""
        cap = (n + self.world - 1) // self.world
        pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        flat = torch.empty(self.world * pad.numel(), dtype=pad.dtype, device=pad.device)
        if pad.is_cuda:
            dist.all_gather_into_tensor(flat, pad.reshape(-1), group=self.group)
        else:  # gloo (CPU tests)
            dist.all_gather(list(flat.view(self.world, -1).unbind(0)), pad.reshape(-1),
                            group=self.group)
        parts = flat.view((self.world,) + tuple(pad.shape))
        return torch.cat([parts[r, :self.rows(n, r)[1]] for r in range(self.world)], dim=0)

    def forward(self, packed: "ops.PackedModel", x: torch.Tensor, mode: str, *, total_members: int,
                precision: str = "fp32", **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        if kw.get("masks") is not None:
            raise ValueError("NShard: injected dropout masks are indexed by global row; use KShard")
        n = x.shape[0]
        begin, count = self.rows(n)
        if mode == "mc_dropout":
            kw["row_base"] = begin
        if kw.get("score_floor") is not None:
            kw["score_floor"] = kw["score_floor"][begin:begin + count]
        if count > 0:
            first, second = packed.forward(x[begin:begin + count], mode,
                                           total_members=total_members, precision=precision, **kw)
            local = torch.stack([first, second], dim=1)              # [count, 2, d_out]
        else:
            local = torch.zeros((0, 2, packed.d_out), dtype=torch.float32, device=x.device)
        full = self.gather_rows(local, n)
        return full[:, 0].contiguous(), full[:, 1].contiguous()


# ------------------------------------------------------------------------------------------------
# Sharded distribution metrics (BASELINE.json configs[4]: 100 M scores over 8 GPUs)
# ------------------------------------------------------------------------------------------------
# Every rank holds an arbitrary shard of the ID scores ``u`` and of the OOD scores ``v``.
#
# KDE-JS (evaluation.py:268-276): all-gather of 2 x 5 float64 shard statistics -> global
# min/max/Scott bandwidth (Chan merge) -> every rank adds its shard's kernel sums to the full
# grid -> ONE all-reduce of 2 x grid_pts float64 -> Jensen-Shannon distance (computed redundantly
# on every rank).
#
# Wasserstein (evaluation.py:182), binned method (default): every rank adds its shards to four
# per-key-bin tables (count and integer offset sum of u and of v) in one pass, ONE all-reduce of
# 4 x 16384 int64 completes them, and every bin on which F_u - F_v keeps one sign is integrated
# from the tables alone.  Only the values of the remaining (ambiguous) bins cross the wire (one
# variable-size all-gather per sample) and are integrated exactly on every rank.  If more than
# half of the values are ambiguous, or a value is inf/NaN, the sample sort below runs instead.
#
# Wasserstein, sort method: sample sort.  One all-reduce of a 2 x 16384-bin key histogram
# picks value-range splitters that balance u+v over the ranks; ONE all-to-all per sample moves
# every value to the rank that owns its range; each rank sorts and integrates |F_u - F_v| over
# its range with the global CDF offsets; an all-gather of 4 float64 per rank adds the partial
# integrals and the terms that straddle two ranges.

class CudaMetricBackend:
    """The product backend: every step is a kernel in libnnueehcs_b200.so (``ops``)."""
    key_bins = staticmethod(ops.key_bins)
    sample_stats = staticmethod(ops.sample_stats)
    kde_grid_accumulate = staticmethod(ops.kde_grid_accumulate)
    jsd_from_grids = staticmethod(ops.jsd_from_grids)
    key_histogram = staticmethod(ops.key_histogram)
    partition_by_bin = staticmethod(ops.partition_by_bin)
    wasserstein_1d_range = staticmethod(ops.wasserstein_1d_range)
    bin_moments = staticmethod(ops.bin_moments)
    wasserstein_from_bins = staticmethod(ops.wasserstein_from_bins)
    compact_flagged = staticmethod(ops.compact_flagged)
    wasserstein_ambiguous = staticmethod(ops.wasserstein_ambiguous)


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("sharded metrics need an initialised torch.distributed process group")
    return dist.get_rank(group), dist.get_world_size(group)


def _all_gather_f64(vec: torch.Tensor, group) -> torch.Tensor:
    rank, world = _world(group)
    out = torch.empty((world,) + tuple(vec.shape), dtype=vec.dtype, device=vec.device)
    dist.all_gather_into_tensor(out.view(-1), vec.contiguous().view(-1), group=group) \
        if vec.is_cuda else dist.all_gather(list(out.unbind(0)), vec.contiguous(), group=group)
    return out


def merge_stats(stats: torch.Tensor):
    """Chan merge of per-shard rows (n, min, max, mean, M2) -> (n, min, max, mean, M2)."""
    n, mn, mx, mean, m2 = 0.0, float("inf"), float("-inf"), 0.0, 0.0
    for row in stats.tolist():
        nb, mnb, mxb, meanb, m2b = row
        if nb <= 0:
            continue
        mn, mx = min(mn, mnb), max(mx, mxb)
        tot = n + nb
        delta = meanb - mean
        mean = mean + delta * nb / tot
        m2 = m2 + m2b + delta * delta * n * nb / tot
        n = tot
    return n, mn, mx, mean, m2


def kde_jsd_sharded(u_local: torch.Tensor, v_local: torch.Tensor, num_points: int = 20000,
                    group=None, backend=CudaMetricBackend) -> float:
    """``JensenShannonEvaluation.pdf_jsd`` over samples sharded across the ranks of ``group``."""
    rank, world = _world(group)
    dev = u_local.device
    rows = []
    for x in (u_local, v_local):
        x = x.reshape(-1)
        if x.numel():
            mn, mx, mean, m2 = backend.sample_stats(x)
            rows.append([float(x.numel()), mn, mx, mean, m2])
        else:
            rows.append([0.0, 0.0, 0.0, 0.0, 0.0])
    gathered = _all_gather_f64(torch.tensor(rows, dtype=torch.float64, device=dev), group)
    nu, mnu, mxu, _, m2u = merge_stats(gathered[:, 0])
    nv, mnv, mxv, _, m2v = merge_stats(gathered[:, 1])
    if nu < 2 or nv < 2:
        raise ValueError("kde_jsd_sharded: each sample needs at least 2 values")
    # scipy.stats.gaussian_kde: Scott factor n^(-1/5) times the unbiased standard deviation
    h_u = (m2u / (nu - 1.0)) ** 0.5 * nu ** -0.2
    h_v = (m2v / (nv - 1.0)) ** 0.5 * nv ** -0.2
    lo, hi = min(mnu, mnv), max(mxu, mxv)
    grids = torch.zeros((2, num_points), dtype=torch.float64, device=dev)
    if u_local.numel():
        backend.kde_grid_accumulate(u_local.reshape(-1), lo, hi, h_u, grids[0])
    if v_local.numel():
        backend.kde_grid_accumulate(v_local.reshape(-1), lo, hi, h_v, grids[1])
    dist.all_reduce(grids, op=dist.ReduceOp.SUM, group=group)
    return backend.jsd_from_grids(grids)


def _all_to_all(send: torch.Tensor, in_splits: List[int], out_splits: List[int], group
                ) -> torch.Tensor:
    recv = torch.empty(sum(out_splits), dtype=send.dtype, device=send.device)
    if send.is_cuda:
        dist.all_to_all_single(recv, send, out_splits, in_splits, group=group)
        return recv
    # gloo (CPU tests) has no all-to-all: pairwise exchange
    rank, world = _world(group)
    s_off = [0]
    for c in in_splits:
        s_off.append(s_off[-1] + c)
    r_off = [0]
    for c in out_splits:
        r_off.append(r_off[-1] + c)
    recv[r_off[rank]:r_off[rank + 1]] = send[s_off[rank]:s_off[rank + 1]]
    reqs = []
    for peer in range(world):
        if peer == rank:
            continue
        if in_splits[peer]:
            reqs.append(dist.isend(send[s_off[peer]:s_off[peer + 1]].contiguous(), peer, group=group))
        if out_splits[peer]:
            reqs.append(dist.irecv(recv[r_off[peer]:r_off[peer + 1]], peer, group=group))
    for r in reqs:
        r.wait()
    return recv


def choose_bin_owners(hist_total, world: int):
    """bin -> owning rank so that the ranks' value ranges are contiguous, ordered and balanced.
    ``hist_total``: 1-D integer array (u + v counts per key bin, all ranks)."""
    import numpy as np
    cum = np.cumsum(np.asarray(hist_total, dtype=np.int64))
    total = int(cum[-1])
    targets = [(p + 1) * total / world for p in range(world - 1)]
    ends = np.searchsorted(cum, targets, side="left")          # last bin of parts 0..world-2
    owners = np.searchsorted(ends, np.arange(cum.size), side="left")
    return owners.astype(np.uint8)


def _all_gather_var(x: torch.Tensor, group) -> torch.Tensor:
    """Concatenation of every rank's 1-D float32 ``x`` (sizes differ)."""
    rank, world = _world(group)
    sizes = _all_gather_f64(torch.tensor([x.numel()], dtype=torch.int64, device=x.device), group)
    sizes = [int(c) for c in sizes.view(-1).tolist()]
    cap = max(sizes)
    if cap == 0:
        return x[:0]
    pad = torch.zeros(cap, dtype=x.dtype, device=x.device)
    pad[:x.numel()] = x
    got = _all_gather_f64(pad, group)
    return torch.cat([got[r, :sizes[r]] for r in range(world)])


def wasserstein_1d_sharded(u_local: torch.Tensor, v_local: torch.Tensor, group=None,
                           backend=CudaMetricBackend, method: str = "auto",
                           info: Optional[dict] = None) -> float:
    """``scipy.stats.wasserstein_distance(u, v)`` over samples sharded across ``group``.
    ``method``: 'auto' | 'binned' | 'sort' (module comment above); ``info`` (optional dict) receives
    the method used and how many values were exchanged."""
    if method not in ("auto", "binned", "sort"):
        raise ValueError(f"unknown Wasserstein method {method!r} (auto, binned, sort)")
    rank, world = _world(group)
    dev = u_local.device
    u_local, v_local = u_local.reshape(-1), v_local.reshape(-1)
    if method != "sort":
        tables = torch.zeros((4, backend.key_bins()), dtype=torch.int64, device=dev)
        backend.bin_moments(u_local, tables, 0)
        backend.bin_moments(v_local, tables, 2)
        dist.all_reduce(tables, op=dist.ReduceOp.SUM, group=group)
        totals = tables[0::2].sum(dim=1).tolist()
        nu_total, nv_total = int(totals[0]), int(totals[1])
        if nu_total == 0 or nv_total == 0:
            raise ValueError("Distribution can't be empty.")
        r = backend.wasserstein_from_bins(tables, nu_total, nv_total)
        amb = r["amb_u"] + r["amb_v"]
        if r["nonfinite"] == 0 and (method == "binned" or amb <= (nu_total + nv_total) // 2):
            if info is not None:
                info.update(method="binned", exchanged_values=amb)
            if amb == 0:
                return r["resolved"]
            all_u = _all_gather_var(backend.compact_flagged(u_local, r["flags"]), group)
            all_v = _all_gather_var(backend.compact_flagged(v_local, r["flags"]), group)
            return r["resolved"] + backend.wasserstein_ambiguous(all_u, all_v, tables, nu_total,
                                                                 nv_total)
    return _wasserstein_1d_sample_sort(u_local, v_local, group, backend, info)


def _wasserstein_1d_sample_sort(u_local, v_local, group, backend, info) -> float:
    import numpy as np
    rank, world = _world(group)
    if world > 64:
        raise ValueError("wasserstein_1d_sharded supports at most 64 ranks")
    dev = u_local.device
    local_hist = torch.stack([backend.key_histogram(u_local), backend.key_histogram(v_local)])
    global_hist = local_hist.clone()
    dist.all_reduce(global_hist, op=dist.ReduceOp.SUM, group=group)
    gh = global_hist.cpu().numpy()
    lh = local_hist.cpu().numpy()
    nu_total, nv_total = int(gh[0].sum()), int(gh[1].sum())
    if nu_total == 0 or nv_total == 0:
        raise ValueError("Distribution can't be empty.")
    owners = choose_bin_owners(gh[0] + gh[1], world)
    owners_t = torch.from_numpy(owners).to(dev)

    def route(x, hist_row):
        send_counts = [int(hist_row[owners == p].sum()) for p in range(world)]
        send = backend.partition_by_bin(x, owners_t, send_counts) if x.numel() else x
        cnt = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        all_cnt = _all_gather_f64(cnt, group)                  # [src, dst]
        recv_counts = [int(c) for c in all_cnt[:, rank].tolist()]
        return _all_to_all(send, send_counts, recv_counts, group)

    mine_u = route(u_local, lh[0])
    mine_v = route(v_local, lh[1])
    if info is not None:
        info.update(method="sort", exchanged_values=nu_total + nv_total)
    below = owners < rank
    u_below, v_below = int(gh[0][below].sum()), int(gh[1][below].sum())
    if mine_u.numel() + mine_v.numel() > 0:
        part, first, last = backend.wasserstein_1d_range(mine_u, mine_v, u_below, v_below,
                                                         nu_total, nv_total)
        row = [part, first, last, float(mine_u.numel() + mine_v.numel())]
    else:
        row = [0.0, 0.0, 0.0, 0.0]
    rows = _all_gather_f64(torch.tensor(row, dtype=torch.float64, device=dev), group).tolist()
    total = 0.0
    prev = None        # (last value, cdf_u, cdf_v) of the previous non-empty range
    for g, (part, first, last, cnt) in enumerate(rows):
        if cnt <= 0:
            continue
        if prev is not None:
            total += abs(prev[1] - prev[2]) * (first - prev[0])
        total += part
        upto = owners <= g
        prev = (last, float(gh[0][upto].sum()) / nu_total, float(gh[1][upto].sum()) / nv_total)
    return total
