#!/usr/bin/env python
"""KDE-JS only, BASELINE configs[4] shape on one GPU -- the command the ncu launch list of the
moment method is taken with (tools/bench_metrics.py times it)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nnueehcs_b200 import ops  # noqa: E402
from tools.bench_metrics import gamma_scores  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
dev = torch.device("cuda:0")
u, v = gamma_scores(n, 2, 0.05, 0, dev), gamma_scores(n, 3, 0.08, 1, dev)
for method in ("auto", "auto", "auto"):
    print(method, ops.kde_jsd_info(u, v, 20000, method))
