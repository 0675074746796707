#!/usr/bin/env python
"""Throughput of the ID-vs-OOD metric kernels (BASELINE.json configs[4], single-GPU slice).

    python tools/bench_metrics.py [--n 50000000] [--steps 5] [--cpu]

ID scores ~ Gamma(2, 0.05), OOD scores ~ Gamma(3, 0.08), float32, resident in
HBM.  Prints one JSON line per metric: values/s, milliseconds, and the HBM roofline fraction
computed from the ALGORITHMIC bytes (4 B per input value read once) next to the measured peak.
``--cpu`` also times the oracle (scipy's algorithm restated in numpy) on a bounded sample.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json"
    return 6650.0, "B200_PROFILING.md fallback"


def gamma_scores(n, shape, scale, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    # Gamma(k, theta) for integer k = sum of k exponentials
    u = torch.rand((shape, n), generator=g, device=dev, dtype=torch.float32).clamp_min_(1e-12)
    return (-torch.log(u)).sum(0).mul_(scale).contiguous()


def time_gpu(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    return sum(ms) / len(ms), min(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50_000_000, help="values per sample (ID and OOD each)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", type=int, default=20000)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    from nnueehcs_b200 import ops
    dev = torch.device("cuda:0")
    u = gamma_scores(args.n, 2, 0.05, 0, dev)
    v = gamma_scores(args.n, 3, 0.08, 1, dev)
    hbm, src = peaks()
    total = 2 * args.n
    same = gamma_scores(args.n, 2, 0.05, 2, dev)   # a second ID sample: every bin is ambiguous
    for name, fn in (("wasserstein_1d", lambda: ops.wasserstein_1d(u, v)),
                     ("wasserstein_1d[sort]", lambda: ops.wasserstein_1d(u, v, "sort")),
                     ("wasserstein_1d[same distribution, auto]", lambda: ops.wasserstein_1d(u, same)),
                     ("kde_jsd", lambda: ops.kde_jsd(u, v, args.grid)),
                     ("kde_jsd[window]", lambda: ops.kde_jsd(u, v, args.grid, "window"))):
        ops.reset_launch_count()
        val = fn()
        launches = ops.launch_count()
        mean_ms, best_ms = time_gpu(fn, args.steps)
        gbs = total * 4 / (mean_ms * 1e-3) / 1e9
        line = {"metric": name, "result": val, "values": total, "ms": mean_ms, "best_ms": best_ms,
                "values_per_s": total / (mean_ms * 1e-3), "gpu_launches_per_call": int(launches),
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                             "frac": gbs / hbm, "algorithmic_bytes": total * 4, "peak_source": src}}
        if name.startswith("wasserstein"):
            info = ops.wasserstein_1d_info(u, same if "same" in name else v,
                                           "sort" if "[sort]" in name else "auto")
            line["method"] = info["method"]
            line["values_sorted"] = info["sorted_u"] + info["sorted_v"]
        if name.startswith("kde_jsd"):
            line["method"] = ops.kde_jsd_info(u, v, args.grid,
                                              "window" if "[window]" in name else "auto")["method"]
            line["gaussian_terms_equivalent_per_s"] = total * args.grid / (mean_ms * 1e-3)
        print(json.dumps(line), flush=True)
    # device time per call of the two single-launch metrics, host out of the loop (enqueue / finish)
    for name, enq in (("wasserstein_1d[enqueued]", lambda: ops.wasserstein_1d_async(u, v)),
                      ("kde_jsd[enqueued]", lambda: ops.kde_jsd_async(u, v, args.grid))):
        [p.result() for p in [enq() for _ in range(2)]]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pend = [enq() for _ in range(args.steps)]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        vals = [p.result() for p in pend]
        gbs = total * 4 / (ms * 1e-3) / 1e9
        print(json.dumps({"metric": name, "result": vals[0], "values": total, "ms": ms,
                          "values_per_s": total / (ms * 1e-3),
                          "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                                       "frac": gbs / hbm, "algorithmic_bytes": total * 4,
                                       "peak_source": src},
                          "timing": f"{args.steps} calls enqueued back to back, one synchronisation"}),
              flush=True)
    # the radix sort on its own: 36 B of HBM traffic per key is what a 4-pass LSD sort with a
    # separate histogram read moves (4 x (4 read + 4 write) + 4 x 4 upsweep read = 48 here)
    lib = ops._lib.load()
    out = torch.empty_like(u)
    wsb = int(lib.uq_sort_workspace_bytes(u.numel()))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream

    def sort_only():
        ops._lib.check(lib.uq_sort_f32(u.data_ptr(), u.numel(), out.data_ptr(), ws.data_ptr(), wsb, st))
    mean_ms, best_ms = time_gpu(sort_only, args.steps)
    ok = bool((out[1:] >= out[:-1]).all().item())
    print(json.dumps({"metric": "radix_sort_f32", "keys": u.numel(), "ms": mean_ms, "best_ms": best_ms,
                      "keys_per_s": u.numel() / (mean_ms * 1e-3), "sorted": ok,
                      "includes": "one device-to-device copy of the input (the sort is in place)",
                      "items_per_lane": os.environ.get("UQ_SORT_ITEMS", "auto")}), flush=True)
    # KDEMLPModel's input-density score: 1 M queries x 100 k fitted rows, d = 5
    g = torch.Generator(device=dev).manual_seed(7)
    fit = torch.rand(100_000, 5, device=dev, generator=g)
    xq = torch.rand(1 << 20, 5, device=dev, generator=g)
    h = ops.kde_scott_bandwidth(*fit.shape)
    mean_ms, best_ms = time_gpu(lambda: ops.kde_density(fit, xq, h), max(2, args.steps // 2))
    pairs = fit.shape[0] * xq.shape[0]
    print(json.dumps({"metric": "kde_density", "queries": xq.shape[0], "fitted_rows": fit.shape[0],
                      "d": 5, "ms": mean_ms, "best_ms": best_ms,
                      "gaussian_terms_per_s": pairs / (mean_ms * 1e-3),
                      "bound": "FP32 + MUFU pipes (N x M terms of 2d + 2 FP32 ops and one ex2)"}),
          flush=True)
    if args.cpu:
        from oracle import metrics_oracle
        un, vn = u[:5_000_000].cpu().numpy(), v[:5_000_000].cpu().numpy()
        t0 = time.perf_counter()
        w = metrics_oracle.wasserstein_1d(un, vn)
        dt = time.perf_counter() - t0
        print(json.dumps({"cpu_baseline": "wasserstein_1d (oracle port of scipy _cdf_distance)",
                          "values": int(un.size + vn.size), "s": dt,
                          "values_per_s": (un.size + vn.size) / dt, "result": w,
                          "cores": torch.get_num_threads()}), flush=True)
        us, vs = un[:20000], vn[:20000]
        t0 = time.perf_counter()
        j = metrics_oracle.pdf_jsd(us, vs, args.grid)
        dt = time.perf_counter() - t0
        print(json.dumps({"cpu_baseline": "pdf_jsd (oracle port of scipy gaussian_kde + jensenshannon)",
                          "values": int(us.size + vs.size), "s": dt,
                          "values_per_s": (us.size + vs.size) / dt, "result": j,
                          "gaussian_terms_per_s": (us.size + vs.size) * args.grid / dt}), flush=True)


if __name__ == "__main__":
    main()
