#!/usr/bin/env python
"""KDE input-density score only (1 M queries x 100 k fitted rows, d = 5): the command the ncu
capture of kde_density_kernel is taken with (tools/bench_metrics.py times it)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nnueehcs_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
fit = torch.rand(100_000, 5, device=dev, generator=g)
x = torch.rand(1 << 20, 5, device=dev, generator=g)
for _ in range(2):
    d = ops.kde_density(fit, x, ops.kde_scott_bandwidth(*fit.shape))
print(float(d.mean()))
