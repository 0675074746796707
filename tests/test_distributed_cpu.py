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

from nnueehcs_b200.distributed import KShard, split_range
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
