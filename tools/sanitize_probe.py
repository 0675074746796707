#!/usr/bin/env python
"""Small inputs through the kernels added or rewritten in round 2 (radix sort, merge-path
integral, far-query KDE density, enqueue / finish metrics), each checked against numpy / the
oracle; prints OK at the end.  Sized so that it can also run under compute-sanitizer where that
is available (it is closed on the build pool's GPU boxes):

    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py --small
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnueehcs_b200 import ops  # noqa: E402
from oracle import metrics_oracle  # noqa: E402

small = "--small" in sys.argv
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def sort_ref(x):
    b = x.view(np.uint32)
    k = np.where(b >> 31, ~b, b | np.uint32(0x80000000)).astype(np.uint32)
    k.sort()
    return np.where(k >> 31, k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32)


for n in ((1, 33, 4097, 8193, 70_001) if small else (1, 33, 4097, 8193, 70_001, 2_424_833 + 8192 + 5)):
    x = rng.standard_normal(n).astype(np.float32)
    got = ops.sort_f32(torch.from_numpy(x).to(dev)).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), sort_ref(x)), n
    if n > 8:
        got = ops.sort_f32(torch.from_numpy(x).to(dev)[1:]).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), sort_ref(x[1:].copy())), n
for nu, nv in ((1, 1), (5, 3), (4096, 4097), (70_001, 30_003)):
    u = rng.gamma(2.0, 0.05, nu).astype(np.float32)
    v = rng.gamma(3.0, 0.08, nv).astype(np.float32)
    ref = metrics_oracle.wasserstein_1d(u, v)
    ud, vd = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
    for m in ("sort", "binned", "auto"):
        got = ops.wasserstein_1d(ud, vd, m)
        assert abs(got - ref) <= 1e-11 * abs(ref) + 1e-15, (nu, nv, m, got, ref)
    assert ops.wasserstein_1d_async(ud, vd).result() == ops.wasserstein_1d(ud, vd)
    if nu >= 2 and nv >= 2:
        assert ops.kde_jsd_async(ud, vd, 500).result() == ops.kde_jsd(ud, vd, 500)
s = ops.score_metrics(torch.from_numpy(rng.gamma(2.0, 0.05, 9001).astype(np.float32)).to(dev),
                      torch.from_numpy(rng.gamma(3.0, 0.08, 7003).astype(np.float32)).to(dev))
assert 0.5 < s["auroc"] < 1.0
fit = rng.random((700, 5)).astype(np.float32)
xq = np.concatenate([rng.random((64, 5)), 1.0 + rng.random((64, 5)) * 8.0]).astype(np.float32)
h = ops.kde_scott_bandwidth(*fit.shape)
got = ops.kde_density(torch.from_numpy(fit).to(dev), torch.from_numpy(xq).to(dev), h).cpu().numpy()
ref = metrics_oracle.kde_neg_density(fit, xq, h)
assert np.allclose(got, ref, rtol=2e-5, atol=1e-300)
torch.cuda.synchronize()
print("OK")
