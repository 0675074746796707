#!/usr/bin/env python
"""Host-side cost of one metric call: tiny inputs, so the kernels are a few microseconds and the
rest is Python + driver + the final synchronisation."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nnueehcs_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
u = torch.rand(4096, device=dev)
v = torch.rand(4096, device=dev) + 0.3
lib = _lib.load()


def timeit(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


wsb = int(lib.uq_wasserstein_workspace_bytes(u.numel(), v.numel()))
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
out = C.c_double()
st = torch.cuda.current_stream(dev).cuda_stream


def direct_w():
    lib.uq_wasserstein_1d_ex(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(), 0, C.byref(out), None,
                             ws.data_ptr(), wsb, st)


kwsb = int(lib.uq_kde_jsd_workspace_bytes(u.numel(), v.numel(), 2000))
kws = torch.empty(kwsb, dtype=torch.uint8, device=dev)


def direct_k():
    lib.uq_kde_jsd_ex(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(), 2000, 0, C.byref(out), None,
                      kws.data_ptr(), kwsb, st)


print("wasserstein_1d  public op  : %.1f us/call" % timeit(lambda: ops.wasserstein_1d(u, v)))
print("wasserstein_1d  C ABI only : %.1f us/call" % timeit(direct_w))
print("kde_jsd         public op  : %.1f us/call" % timeit(lambda: ops.kde_jsd(u, v, 2000)))
print("kde_jsd         C ABI only : %.1f us/call" % timeit(direct_k))
print("torch.empty(800 MB) cached : %.1f us" % timeit(lambda: torch.empty(800_000_000, dtype=torch.uint8, device=dev)))
print("empty kernel + sync        : %.1f us" % timeit(lambda: (u.add_(0), torch.cuda.synchronize())))
print("wasserstein_1d  torch.ops  : %.1f us/call" % timeit(lambda: torch.ops.nnueehcs_b200.wasserstein_1d(u, v, 0)))
print("wasserstein_1d  _op direct : %.1f us/call" % timeit(lambda: ops._op_wasserstein_1d(u, v, 0)))
# phase times of the single-launch KDE-JS at BASELINE configs[4] size
n = 50_000_000
g = torch.Generator(device=dev).manual_seed(0)
big_u = (-torch.log(torch.rand((2, n), generator=g, device=dev).clamp_min_(1e-12))).sum(0).mul_(0.05)
big_v = (-torch.log(torch.rand((3, n), generator=g, device=dev).clamp_min_(1e-12))).sum(0).mul_(0.08)
for _ in range(3):
    ops.kde_jsd(big_u, big_v, 20000)
ph = (C.c_double * 5)()
lib.uq_kde_jsd_phase_us(ph)
print("kde_jsd 50M+50M phases (us): stats %.1f, fine-bin pass %.1f, fold %.1f, grid %.1f, js %.1f" % tuple(ph))
ops.kde_jsd(u, v, 2000)
lib.uq_kde_jsd_phase_us(ph)
print("kde_jsd 4096+4096 phases (us): stats %.1f, fine-bin pass %.1f, fold %.1f, grid %.1f, js %.1f" % tuple(ph))
