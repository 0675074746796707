"""Timing of uq_score_metrics (AUROC, TNR@TPR, percentiles, classifier rates) at 50 M + 50 M."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nnueehcs_b200 import ops
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
def gamma(shape, scale, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    u = torch.rand((shape, n), generator=g, device=dev).clamp_min_(1e-12)
    return (-torch.log(u)).sum(0).mul_(scale).contiguous()
u, v = gamma(2, 0.05, 0), gamma(3, 0.08, 1)
for _ in range(2):
    r = ops.score_metrics(u, v)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    r = ops.score_metrics(u, v)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"metric": "score_metrics", "values": 2 * n, "ms": e0.elapsed_time(e1) / 5,
                  "auroc": r["auroc"], "tnr_at_tpr": r["tnr_at_tpr"]}))
