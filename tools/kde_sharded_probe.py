"""Phase timing of distributed.kde_jsd_sharded / wasserstein_1d_sharded (torchrun)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", device_id=dev)
from nnueehcs_b200 import distributed as nd, ops
def gamma(n, shape, scale, seed):
    g = torch.Generator(device=dev).manual_seed(seed * 1000 + rank)
    u = torch.rand((shape, n), generator=g, device=dev).clamp_min_(1e-12)
    return (-torch.log(u)).sum(0).mul_(scale).contiguous()
n = 50_000_000 // world
u, v = gamma(n, 2, 0.05, 1), gamma(n, 3, 0.08, 2)
class Timed:
    def __init__(self): self.t = {}
    def wrap(self, name, fn):
        def f(*a, **k):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize(); self.t[name] = self.t.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return r
        return staticmethod(f)
T = Timed()
class B(nd.CudaMetricBackend): pass
for name in ("sample_stats", "kde_grid_accumulate", "jsd_from_grids", "key_histogram", "partition_by_bin", "wasserstein_1d_range"):
    setattr(B, name, T.wrap(name, getattr(ops, name)))
orig_ar, orig_ag, orig_a2a = dist.all_reduce, dist.all_gather_into_tensor, dist.all_to_all_single
dist.all_reduce = T.wrap("all_reduce", orig_ar).__func__
dist.all_gather_into_tensor = T.wrap("all_gather", orig_ag).__func__
dist.all_to_all_single = T.wrap("all_to_all", orig_a2a).__func__
for fn, label in ((lambda: nd.kde_jsd_sharded(u, v, 20000, backend=B), "kde"), (lambda: nd.wasserstein_1d_sharded(u, v, backend=B), "wasserstein")):
    fn(); T.t.clear()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); total = (time.perf_counter() - t0) / 3 * 1e3
    if rank == 0:
        print(json.dumps({"what": label, "world": world, "total_ms": round(total, 3), "phases_ms": {k: round(x / 3, 3) for k, x in T.t.items()}}), flush=True)
    T.t.clear()
dist.destroy_process_group()
