"""World-size-2 ``gloo`` test of the K-axis sharding host logic (runs on CPU).

The product's moment merge is a CUDA kernel, so here the per-shard moments and the Chan merge are
computed by the oracle; what is under test is ``nnueehcs_b200.distributed``: the balanced split of
the member axis, the single all-gather exchange and its rank ordering, and that merging the
exchanged shards reproduces the unsharded mean/std of the reference arithmetic.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nnueehcs_b200.distributed import KShard, NShard, split_range
from oracle import uq_oracle
from tests.util import load_golden, nets_from_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden("ensemble_bn.npz")
        k = int(g["k"])  # 5 members over 2 ranks -> 2 + 3
        nets = nets_from_golden(g, k)
        x = torch.from_numpy(g["x"])
        shard = KShard()
        begin, count = shard.split(k)
        assert (begin, count) == split_range(k, world, rank)
        assert sum(shard.counts(k)) == k and shard.counts(k) == [2, 3]
        members = uq_oracle.ensemble_member_outputs(nets[begin:begin + count], x)
        n, mean, m2 = uq_oracle.moments_of(members)
        means, m2s = shard.exchange(mean.float(), m2.float())
        assert means.shape == (world,) + tuple(mean.shape)
        # rank ordering of the gathered slab
        assert torch.equal(means[rank], mean.float())
        acc = None
        for r, c in enumerate(shard.counts(k)):
            part = (torch.full_like(means[r], float(c)).double(), means[r].double(), m2s[r].double())
            acc = part if acc is None else uq_oracle.chan_merge(acc, part)
        m, s = uq_oracle.finalize_std(*acc)
        ref_mean, ref_std = torch.from_numpy(g["mean"]), torch.from_numpy(g["std"])
        assert float((m - ref_mean).abs().max()) <= 1e-5 * float(ref_mean.abs().max())
        assert float((s - ref_std).abs().max()) <= 1e-5 * float(ref_mean.abs().max())
        torch.save({"mean": m, "std": s}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_split_range_is_balanced_and_contiguous():
    for total in (1, 2, 5, 16, 100, 1000):
        for world in (1, 2, 3, 8):
            parts = [split_range(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (b0, c0), (b1, _) in zip(parts[:-1], parts[1:]):
                assert b0 + c0 == b1
            counts = [c for _, c in parts]
            assert max(counts) - min(counts) <= 1


def test_kshard_requires_process_group():
    with pytest.raises(RuntimeError, match="process group"):
        KShard()


def test_gloo_world2_exchange_and_merge(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "rank0.pt")
    b = torch.load(tmp_path / "rank1.pt")
    # every rank ends with the same merged result
    assert torch.equal(a["mean"], b["mean"]) and torch.equal(a["std"], b["std"])


class _OraclePacked:
    """Stand-in for ops.PackedModel on CPU: the oracle's ensemble forward (test infrastructure)."""

    def __init__(self, nets):
        self.nets, self.d_out, self.calls = nets, 1, []

    def forward(self, x, mode, *, total_members, precision="fp32", **kw):
        self.calls.append((tuple(x.shape), mode, total_members, dict(kw)))
        return uq_oracle.ensemble_forward(self.nets, x)


def _nshard_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden("ensemble_bn.npz")
        k = int(g["k"])
        packed = _OraclePacked(nets_from_golden(g, k))
        x = torch.from_numpy(g["x"])[:157]            # 157 rows over 2 ranks -> 79 + 78
        shard = NShard()
        assert shard.rows(157) == split_range(157, world, rank)
        mean, std = shard.forward(packed, x, "ensemble", total_members=k)
        assert packed.calls[0][0] == (shard.rows(157)[1], x.shape[1])   # only its own rows ran
        ref_mean, ref_std = uq_oracle.ensemble_forward(packed.nets, x)
        assert torch.allclose(mean, ref_mean, rtol=0, atol=1e-7)
        assert torch.allclose(std, ref_std, rtol=0, atol=1e-7)
        # MC dropout: Philox is keyed by the GLOBAL row (row_base = first row of the rank's slice);
        # injected masks are refused
        shard.forward(packed, x, "mc_dropout", total_members=k, offset=10)
        assert packed.calls[-1][3]["offset"] == 10
        assert packed.calls[-1][3]["row_base"] == shard.rows(157)[0]
        with pytest.raises(ValueError, match="indexed by global row"):
            shard.forward(packed, x, "mc_dropout", total_members=k, masks=torch.zeros(1))
        # fewer rows than ranks: one rank holds nothing and still takes part in the gather
        tiny_mean, _ = shard.forward(packed, x[:1], "ensemble", total_members=k)
        assert tiny_mean.shape == (1, 1) and torch.allclose(tiny_mean, ref_mean[:1], atol=1e-7)
        torch.save({"mean": mean}, os.path.join(out_dir, f"nrank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sample_axis_shard(tmp_path):
    mp.spawn(_nshard_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "nrank0.pt")
    b = torch.load(tmp_path / "nrank1.pt")
    assert torch.equal(a["mean"], b["mean"])


# ---- reduce-scatter style combine (all-to-all of row slices -> local merge -> all-gather) ----------

class _OracleMomentsPacked:
    """CPU stand-in for ops.PackedModel.forward_into: oracle moments of the member range."""

    def __init__(self, nets):
        self.nets, self.d_out = nets, 1

    def forward_into(self, x, mode, out0, out1, *, total_members, precision="fp32", member_begin=0,
                     member_count=None, output="mean_std", **kw):
        assert output == "moments"
        count = total_members - member_begin if member_count is None else member_count
        members = uq_oracle.ensemble_member_outputs(self.nets[member_begin:member_begin + count], x)
        _, mean, m2 = uq_oracle.moments_of(members)
        out0.copy_(mean.float())
        out1.copy_(m2.float())


def _merge_standin(shards, counts):
    """numpy-style restatement of uq_moments_merge_ex on CPU: shards [S, 2, L] -> [2, L]."""
    acc = None
    for s, c in enumerate(counts):
        if c <= 0:
            continue
        part = (torch.full_like(shards[s, 0], float(c)).double(), shards[s, 0].double(),
                shards[s, 1].double())
        acc = part if acc is None else uq_oracle.chan_merge(acc, part)
    m, sd = uq_oracle.finalize_std(*acc)
    return torch.stack([m.float(), sd.float()])


def _combine_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden("ensemble_bn.npz")
        k = int(g["k"])
        nets = nets_from_golden(g, k)
        packed = _OracleMomentsPacked(nets)
        x = torch.from_numpy(g["x"])[:157]        # odd row count: the last slice is padded
        shard = KShard()
        assert shard.slice_len(157) == 79 and shard.new_slab(157, "cpu").shape == (2, 158)
        mean, std = shard.forward(packed, x, "ensemble", total_members=k, merge=_merge_standin)
        ref_mean, ref_std = uq_oracle.ensemble_forward(nets, x)
        assert mean.shape == ref_mean.shape
        assert torch.allclose(mean, ref_mean, rtol=0, atol=2e-7)
        assert torch.allclose(std, ref_std, rtol=0, atol=2e-7)
        # members owned per rank (weak-scaling ensembles): rank r owns nets[2r : 2r + 2]
        own = _OracleMomentsPacked(nets[2 * rank:2 * rank + 2])
        m2_, s2_ = shard.forward_owned(own, x, "ensemble", local_members=2, merge=_merge_standin)
        r_mean, r_std = uq_oracle.ensemble_forward(nets[:4], x)
        assert torch.allclose(m2_, r_mean, rtol=0, atol=2e-7)
        assert torch.allclose(s2_, r_std, rtol=0, atol=2e-7)
        # more ranks than members: rank 1 contributes an empty shard (count 0) and still takes part
        one = _OracleMomentsPacked(nets[:1])
        m1, s1 = shard.forward(one, x, "ensemble", total_members=1, merge=_merge_standin)
        assert torch.allclose(m1, uq_oracle.ensemble_member_outputs(nets[:1], x)[0], atol=2e-7)
        assert torch.isnan(s1).all()              # unbiased std of a single member, like torch.std
        # MC-dropout seeds: every rank keys Philox with rank 0's seed, whatever its own generator holds
        from nnueehcs_b200.models import MCDropoutModel
        mc = MCDropoutModel(nets[0], num_samples=4, dropout_percent=0.2)
        mc.uq_shard = shard
        assert mc._shared_seed(1000 + rank, "cpu") == 1000
        torch.save({"mean": mean, "std": std}, os.path.join(out_dir, f"crank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_reduce_scatter_combine(tmp_path):
    mp.spawn(_combine_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "crank0.pt")
    b = torch.load(tmp_path / "crank1.pt")
    assert torch.equal(a["mean"], b["mean"]) and torch.equal(a["std"], b["std"])
