"""Host-side cost of one wrapper call on a small MC-dropout job (BASELINE configs[0]): cProfile of
model(x, return_ue=True) with x resident on the device, plus the wall time per call."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "mcdropout100_binomial_10k"
model = bench.build_model(wl).to("cuda")
model.uq_precision = "bf16"
mode, d_in, widths, d_out, k, n, p = bench.WORKLOADS[wl]
x = torch.rand(min(n, 10000), d_in, device="cuda")
with torch.no_grad():
    for _ in range(20):
        model(x, return_ue=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        model(x, return_ue=True)
    t_enq = (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 200
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        model(x, return_ue=True)
    pr.disable()
    torch.cuda.synchronize()
print(f"{wl}: host time to enqueue one call {t_enq * 1e6:.1f} us, per call incl. GPU {t_all * 1e6:.1f} us")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
print(s.getvalue()[:4000])
