"""World-size-2 ``gloo`` tests of the sharded metric host logic (CPU).

The per-rank CUDA steps are replaced by their numpy stand-ins (``oracle.metrics_oracle``);
what is under test is ``nnueehcs_b200.distributed``: splitter choice, routing counts, the
all-to-all, CDF offsets, the terms that straddle two value ranges, and the statistics merge
behind the KDE bandwidth -- against the unsharded oracle (scipy's algorithm) on the same data.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nnueehcs_b200 import distributed as nd
from oracle import metrics_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _samples(case):
    rng = np.random.default_rng(5)
    if case == "gamma":
        u = rng.gamma(2.0, 0.05, 6000).astype(np.float32)
        v = rng.gamma(3.0, 0.08, 4500).astype(np.float32)
    elif case == "ties":           # heavy ties, identical values in both samples, negatives
        u = rng.integers(-3, 4, 3000).astype(np.float32) * 0.5
        v = rng.integers(-1, 6, 2000).astype(np.float32) * 0.5
    elif case == "crossing":       # CDFs cross: a few ambiguous bins travel, the rest resolves
        u = rng.normal(0.0, 1.0, 5000).astype(np.float32)
        v = rng.normal(0.3, 2.0, 3500).astype(np.float32)
    else:                          # disjoint supports: one rank's range may hold a single sample
        u = rng.uniform(0.0, 1.0, 2500).astype(np.float32)
        v = rng.uniform(5.0, 6.0, 2500).astype(np.float32)
    return u, v


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        u, v = _samples(case)
        # uneven, interleaved shards
        ul = torch.from_numpy(u[rank::world].copy()) if rank else torch.from_numpy(u[0::world][:-7].copy())
        if rank == world - 1:
            ul = torch.cat([ul, torch.from_numpy(u[0::world][-7:].copy())])
        vl = torch.from_numpy(v[rank::world].copy())
        be = metrics_oracle.NumpyShardBackend
        info = {m: {} for m in ("auto", "binned", "sort")}
        w = {m: nd.wasserstein_1d_sharded(ul, vl, backend=be, method=m, info=info[m])
             for m in info}
        j = nd.kde_jsd_sharded(ul, vl, 512, backend=be)
        torch.save({"w": w, "j": j, "info": info}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["gamma", "ties", "crossing", "disjoint"])
def test_sharded_metrics_match_unsharded_oracle(tmp_path, case):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    u, v = _samples(case)
    w_ref = metrics_oracle.wasserstein_1d(u, v)
    j_ref = metrics_oracle.pdf_jsd(u, v, 512)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"), weights_only=False)
        for m, w in got["w"].items():
            assert abs(w - w_ref) <= 1e-12 * max(1.0, abs(w_ref)), (case, m, w, w_ref)
        info = got["info"]
        assert info["sort"]["method"] == "sort" and info["binned"]["method"] == "binned"
        assert info["sort"]["exchanged_values"] == u.size + v.size
        if case in ("gamma", "disjoint"):   # every bin resolves: nothing but the tables is exchanged
            assert info["auto"] == {"method": "binned", "exchanged_values": 0}
        if case == "crossing":
            assert info["auto"]["method"] == "binned"
            assert 0 < info["auto"]["exchanged_values"] < (u.size + v.size) // 4
        assert abs(got["j"] - j_ref) <= 5e-8 * max(1.0, abs(j_ref)), (case, got["j"], j_ref)


def test_choose_bin_owners_is_monotone_and_balanced():
    rng = np.random.default_rng(0)
    hist = rng.integers(0, 50, 16384)
    owners = nd.choose_bin_owners(hist, 8)
    assert owners.min() == 0 and owners.max() == 7
    assert np.all(np.diff(owners.astype(int)) >= 0)
    loads = np.array([hist[owners == p].sum() for p in range(8)])
    assert loads.max() <= 1.05 * hist.sum() / 8 + hist.max()


def test_merge_stats_is_chan():
    rng = np.random.default_rng(1)
    a = rng.normal(3.0, 2.0, 1000)
    parts = np.array_split(a, 3)
    rows = torch.tensor([[p.size, p.min(), p.max(), p.mean(), ((p - p.mean()) ** 2).sum()]
                         for p in parts] + [[0, 0, 0, 0, 0]], dtype=torch.float64)
    n, mn, mx, mean, m2 = nd.merge_stats(rows)
    assert n == a.size and mn == a.min() and mx == a.max()
    assert abs(mean - a.mean()) < 1e-12 and abs(m2 - ((a - a.mean()) ** 2).sum()) < 1e-9
