import sys, time, torch
sys.path.insert(0, '/root/repo')
from nnueehcs_b200 import ops
dev = torch.device('cuda:0')
def gamma(n, shape, scale, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    u = torch.rand((shape, n), generator=g, device=dev).clamp_min_(1e-12)
    return (-torch.log(u)).sum(0).mul_(scale).contiguous()
for n in (6_250_000, 50_000_000):
    u, v = gamma(n, 2, 0.05, 1), gamma(n, 3, 0.08, 2)
    def t(fn, reps=3):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3
    st = ops.sample_stats(u)
    mn, mx, mean, m2 = st
    h = (m2 / (n - 1)) ** 0.5 * n ** -0.2
    grids = torch.zeros((2, 20000), dtype=torch.float64, device=dev)
    print(n, 'stats ms', t(lambda: ops.sample_stats(u)),
          'accumulate(u) ms', t(lambda: ops.kde_grid_accumulate(u, mn, mx * 1.5, h, grids[0])),
          'fused kde_jsd ms', t(lambda: ops.kde_jsd(u, v, 20000)),
          'wasserstein ms', t(lambda: ops.wasserstein_1d(u, v)),
          'hist ms', t(lambda: ops.key_histogram(u)))
