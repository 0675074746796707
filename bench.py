#!/usr/bin/env python
"""Benchmark of the UQ inference hot path (BASELINE.json metric: UQ samples x passes / sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Workload at N = 1 (``configs[1]`` of BASELINE.json): binomial-options deep ensemble,
16 members x (5 -> 512 -> 512 -> 512 -> 1, Linear+BatchNorm1d+ReLU), 1 M synthetic samples,
bf16 tcgen05 mode.  One step = one ``model(x, return_ue=True)`` over the whole batch.
At N > 1 (one rank per GPU, launched by torchrun) the member axis is sharded: every rank owns 16
more members (weak scaling: 16 N members in total), reduces them to per-sample (mean, M2) in the
fused kernel, and the shards are combined with one NCCL all-gather + Chan merge per step.

The JSON line follows the driver's contract: ``value`` is whole-job sample.members/s with inputs
resident in HBM; ``e2e`` is the same metric through the wrapper API with pinned HOST buffers
(H2D of x and D2H of mean/std inside the timed region); ``roofline`` is the fused kernel against
the measured bf16 tensor peak; ``cpu_baseline`` is the oracle port (the reference's torch-CPU
arithmetic) timed on a bounded sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (mode, d_in, hidden widths, d_out, members/passes, samples, dropout p)
    "ensemble16x512_1M": ("ensemble", 5, [512, 512, 512], 1, 16, 1 << 20, 0.0),
    "mcdropout100_binomial_10k": ("mc_dropout", 5, [128] * 6, 1, 100, 10000, 0.2),
    "deltauq32_binomial_4M": ("delta_uq", 5, [128] * 6, 1, 32, 1 << 22, 0.0),
    "mcdropout_1000x512_64k": ("mc_dropout", 5, [512] * 7, 1, 1000, 1 << 16, 0.2),
    # configs[3] of BASELINE.json (MC dropout, 1000 passes x 8-layer width-1024 MLP) on a
    # 64 k-sample slice of its 16 M samples, and the same net without dropout
    "mcdropout_1000x1024_64k": ("mc_dropout", 5, [1024] * 7, 1, 1000, 1 << 16, 0.2),
    "mcdropout_1000x1024_1M": ("mc_dropout", 5, [1024] * 7, 1, 1000, 1 << 20, 0.2),
    "ensemble8x1024_256k": ("ensemble", 5, [1024] * 7, 1, 8, 1 << 18, 0.0),
    # the binomial-options surrogate as a 32-member ensemble (same flops as deltauq32_binomial_4M)
    "ensemble32x128_4M": ("ensemble", 5, [128] * 6, 1, 32, 1 << 22, 0.0),
}
DEFAULT_WORKLOAD = "ensemble16x512_1M"
N_ROTATE = 8  # input buffers rotated per step so the working set exceeds the 126 MB L2


def kernel_name(widths, d_out):
    """Which fused kernel the library dispatches to (csrc/mlp_tc.cu:tc_forward)."""
    h = max(widths)
    if h > 512:
        return "uq_mlp_tc3_kernel (CTA pairs, 64 rows per CTA)"
    if h <= 128 and d_out == 1:
        return "uq_mlp_tc4_kernel (CTA pairs, 4 tile slots per CTA)"
    return "uq_mlp_tc2_kernel (CTA pairs)"


def flops_per_unit(d_in, widths, d_out):
    dims = [d_in] + list(widths) + [d_out]
    return 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))


def mlp_arch(d_in, widths, d_out):
    arch, prev = [], d_in
    for w in widths:
        arch += [{"Linear": {"args": [prev, w]}}, {"BatchNorm1d": {"args": [w]}},
                 {"ReLU": {"inplace": True}}]
        prev = w
    arch.append({"Linear": {"args": [prev, d_out]}})
    return arch


def randomise_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)


def build_model(workload, member_offset=0):
    """Random-init model of the named architecture through the builder mirror (the reference's
    own construction path, model_builder.py:219-275); BN running stats randomised (seed 1+i)."""
    from nnueehcs_b200 import model_builder as mb
    mode, d_in, widths, d_out, k, n, p = WORKLOADS[workload]
    if mode == "ensemble":
        arch = mlp_arch(d_in, widths, d_out)
        builder = mb.EnsembleModelBuilder(arch, {"num_models": k})
        model = builder.build()
        if member_offset:  # other ranks own other members: seeds 42 + offset + i
            for i in range(k):
                torch.manual_seed(42 + member_offset + i)
                model.models[i] = mb.build_network(arch)
        for i, net in enumerate(model.models):
            randomise_bn(net, 1 + member_offset + i)
    elif mode == "mc_dropout":
        torch.manual_seed(42)
        model = mb.MCDropoutModelBuilder(mlp_arch(d_in, widths, d_out),
                                         {"num_samples": k, "dropout_percent": p}).build()
        randomise_bn(model.model, 1)
    else:
        torch.manual_seed(42)
        model = mb.DeltaUQMLPModelBuilder(mlp_arch(d_in, widths, d_out),
                                          {"estimator": "std", "num_anchors": k,
                                           "anchored_batch_size": 4096}).build()
        randomise_bn(model.net, 1)
        model.anchors = torch.rand(k, d_in, generator=torch.Generator().manual_seed(2))
    model.eval()
    return model


def synth_x(n, d_in, seed):
    # binomial-options-shaped: 5 features min-max scaled to [0, 1] (data_utils.py:291-296)
    return torch.rand(n, d_in, generator=torch.Generator().manual_seed(seed))


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port (reference arithmetic on torch CPU)
# ------------------------------------------------------------------------------------------------

def cpu_forward_fn(workload, model):
    from oracle import uq_oracle
    mode, d_in, widths, d_out, k, n, p = WORKLOADS[workload]
    if mode == "ensemble":
        nets = list(model.models)
        return lambda x: uq_oracle.ensemble_forward(nets, x)
    if mode == "mc_dropout":
        return lambda x: uq_oracle.mc_dropout_forward(model.model, x, k, p)
    anchors = model.anchors
    return lambda x: uq_oracle.delta_uq_forward(model.net, x, anchors, k)


def time_cpu(workload, model, sample_n, steps, warmup):
    mode, d_in, widths, d_out, k, n, p = WORKLOADS[workload]
    fn = cpu_forward_fn(workload, model)
    x = synth_x(sample_n, d_in, 0)
    for _ in range(warmup):
        fn(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(x)
    dt = (time.perf_counter() - t0) / steps
    return sample_n * k / dt, dt


def cpu_sample_size(workload):
    mode, d_in, widths, d_out, k, n, p = WORKLOADS[workload]
    # a few seconds of CPU work per step (10-30 s for the whole bounded baseline run)
    target_flops = 1.2e12
    s = int(target_flops / (flops_per_unit(d_in, widths, d_out) * k))
    return max(256, min(n, 1 << (s.bit_length() - 1)))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port, kind
    "port") on this box's host cores, same metric/config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # all the host threads it can use: torchrun exports OMP_NUM_THREADS=1 to every rank, which
    # would time the reference on one core (measured: 8.0e4 instead of 7.2e5 sample*members/s)
    try:
        host_threads = len(os.sched_getaffinity(0))
    except AttributeError:
        host_threads = os.cpu_count() or 1
    torch.set_num_threads(max(1, host_threads))
    wl = args.workload
    mode, d_in, widths, d_out, k, n, p = WORKLOADS[wl]
    model = build_model(wl)
    sample_n = cpu_sample_size(wl)
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    # keep the whole run within a few minutes
    probe_rate, probe_dt = time_cpu(wl, model, sample_n, 1, 1)
    steps = max(1, min(steps, int(120.0 / max(probe_dt, 1e-3))))
    rate, dt = time_cpu(wl, model, sample_n, steps, 0)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "uq_sample_passes_per_sec", "value": rate,
        "unit": "sample*members/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(wl, 1, sample_note=f"{sample_n} of {n} samples per step"),
        "cpu_baseline": {"value": rate, "unit": "sample*members/s", "cores": cores,
                         "kind": "port",
                         "sample": f"{sample_n} samples x {k} members per step, {steps} steps, "
                                   f"torch {torch.__version__} CPU, os.cpu_count={os.cpu_count()}"},
        "e2e": {"value": rate, "unit": "sample*members/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def config_dict(wl, n_gpus, sample_note=None):
    mode, d_in, widths, d_out, k, n, p = WORKLOADS[wl]
    cfg = {"workload": wl, "mode": mode, "arch": f"{d_in}->" + "->".join(map(str, widths)) + f"->{d_out}",
           "members_per_gpu": k, "members_total": k * n_gpus if mode == "ensemble" else k,
           "samples": n, "dropout_p": p,
           "parallelism": f"member-axis shards x{n_gpus}, one all-gather of (mean, M2) per step",
           "l2": f"inputs rotated over {N_ROTATE} device buffers (> 126 MB L2 with weights)"}
    if sample_note:
        cfg["cpu_sample"] = sample_note
    return cfg


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]  # samples under load dominate the upper half
        return {"sm_mhz": statistics.median(sm), "sm_mhz_under_load": statistics.median(busy),
                "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"burst": d.get("bf16_tflops", 1590.0), "sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm": d.get("hbm_gbs", 6650.0), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0,
            "source": "B200_PROFILING.md fallback (of fallback)"}


def metric_kernel_lines(dev, peaks, n=50_000_000, steps=5):
    """BASELINE.json's metric also asks for the metric kernels' GB/s: ID-vs-OOD Wasserstein and
    KDE Jensen-Shannon over 50 M + 50 M synthetic uncertainty scores (configs[4] on one GPU),
    resident in HBM, timed with CUDA events around the public op; GB/s = the algorithmic 4 B per
    value over the call time, next to the measured HBM peak."""
    from nnueehcs_b200 import ops

    def gamma(shape, scale, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        u = torch.rand((shape, n), generator=g, device=dev).clamp_min_(1e-12)
        return (-torch.log(u)).sum(0).mul_(scale).contiguous()

    u, v = gamma(2, 0.05, 0), gamma(3, 0.08, 1)
    hbm = peaks.get("hbm")
    out = {}
    for name, fn, info in (
            ("wasserstein_1d", lambda: ops.wasserstein_1d(u, v), lambda: ops.wasserstein_1d_info(u, v)),
            ("kde_jsd", lambda: ops.kde_jsd(u, v, 20000), lambda: ops.kde_jsd_info(u, v, 20000))):
        for _ in range(3):
            val = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.reset_launch_count()
        e0.record()
        for _ in range(steps):
            val = fn()
        e1.record()
        torch.cuda.synchronize()
        launches = ops.launch_count() // steps
        ms = e0.elapsed_time(e1) / steps
        gbs = 2 * n * 4 / (ms * 1e-3) / 1e9
        out[name] = {"values": 2 * n, "ms": ms, "values_per_s": 2 * n / (ms * 1e-3),
                     "algorithmic_GBps": gbs, "hbm_peak_GBps": hbm,
                     "frac_of_hbm_peak": gbs / hbm if hbm else None, "result": val,
                     "method": info()["method"], "gpu_launches_per_call": int(launches),
                     "data": "synthetic Gamma(2, 0.05) vs Gamma(3, 0.08) scores, float32, in HBM"}
    del u, v
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import torch.distributed as dist
    from nnueehcs_b200 import ops
    from nnueehcs_b200.distributed import KShard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference "
                           "for the CPU baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator comes up; stdout
        # carries exactly one JSON line, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    wl = args.workload
    mode, d_in, widths, d_out, k, n, p = WORKLOADS[wl]
    precision = args.precision
    model = build_model(wl, member_offset=rank * k if mode == "ensemble" else 0)
    model.to(dev)
    model.eval()
    model.uq_precision = precision
    shard = KShard() if world > 1 else None
    nets = (list(model.models) if mode == "ensemble" else
            [model.model] if mode == "mc_dropout" else [model.net])
    packed = model._packed(nets, dev)
    if precision == "bf16" and not packed.supports_bf16:
        raise RuntimeError(f"bf16 path unavailable: {packed.bf16_reason}")

    xs_host = [synth_x(n, d_in, s).pin_memory() for s in range(2)]
    xs = [synth_x(n, d_in, s).to(dev) for s in range(N_ROTATE)]
    kw = {}
    if mode == "mc_dropout":
        kw = dict(dropout_p=p, dropout_active=True, seed=1234)
    if mode == "delta_uq":
        kw = dict(anchors=model.anchors.to(dev))

    def step(x):
        if shard is not None and mode == "ensemble":
            return shard.forward_owned(packed, x, mode, local_members=k, precision=precision, **kw)
        if shard is not None:
            return shard.forward(packed, x, mode, total_members=k, precision=precision, **kw)
        return packed.forward(x, mode, total_members=k, precision=precision, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = args.steps, max(3, args.warmup)
    for i in range(warmup):
        out = step(xs[i % N_ROTATE])
    barrier()

    # ---- device-resident timing: CUDA events on the launching (current) stream -----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    ops.reset_launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_begin.record()
    for i in range(steps):
        ev[i][0].record()
        out = step(xs[(warmup + i) % N_ROTATE])
        ev[i][1].record()
    e_end.record()
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = ops.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = e_begin.elapsed_time(e_end)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    units_per_step = n * k * (world if mode == "ensemble" else 1)
    value = units_per_step * steps / (total_ms * 1e-3)

    # ---- end-to-end through the wrapper API with pinned HOST buffers ---------------------------
    model.uq_shard = None
    mean_h = torch.empty((n, d_out), dtype=torch.float32).pin_memory()
    std_h = torch.empty((n, d_out), dtype=torch.float32).pin_memory()

    def e2e_step(xh):
        xd = xh.to(dev, non_blocking=True)
        if shard is not None:
            mean, std = step(xd)
        else:
            with torch.no_grad():
                if mode == "mc_dropout":
                    torch.manual_seed(0)
                mean, std = model(xd, return_ue=True)
        mean_h.copy_(mean, non_blocking=True)
        std_h.copy_(std, non_blocking=True)
        torch.cuda.synchronize()

    e2e_steps = max(3, min(steps, 10))
    for i in range(2):
        e2e_step(xs_host[i % 2])
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(xs_host[i % 2])
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = units_per_step * e2e_steps / e2e_s

    if rank == 0:
        peaks = measured_peaks()
        F = flops_per_unit(d_in, widths, d_out)
        kernel_ms = statistics.mean(step_ms)
        timed_s = total_ms * 1e-3
        peak_kind = "sustained" if timed_s >= 2.0 else "burst"
        peak = peaks[peak_kind]
        # per-GPU rate: an ensemble rank runs its own k members, a K-sharded job k / world of them
        achieved = F * n * (k if mode == "ensemble" else k / world) / (kernel_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(wl, {}).get(precision)
        line = {
            "metric": "uq_sample_passes_per_sec", "value": value, "unit": "sample*members/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
            "higher_is_better": True,
            "scaling": "weak" if mode == "ensemble" else "strong", "vs_baseline": None,
            "dtype": precision if precision == "bf16" else "f32", "data": "synthetic",
            "config": config_dict(wl, world),
            "e2e": {"value": e2e_value, "unit": "sample*members/s",
                    "h2d_bytes_per_step": n * d_in * 4, "d2h_bytes_per_step": 2 * n * d_out * 4,
                    "steps": e2e_steps, "api": "model(x.to(device), return_ue=True) -> pinned host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "peak_source": peaks["source"], "kernel": kernel_name(widths, d_out)
                         if precision == "bf16" else
                         ("uq_mlp_tcx_kernel (fp32 parity: scaled fp16 x 2 split on tcgen05, 3 MMAs "
                          "per K step)" if packed.fp32_on_tensor_cores
                          else "sgemm_tn_kernel (fp32 CUDA cores)"),
                         "flops_per_unit": F, "kernel_ms": kernel_ms,
                         "frac_of_sustained": achieved / peaks["sustained"]},
            "wall_s_timed_region": wall,
        }
        if world == 1 and wl == DEFAULT_WORKLOAD and not args.no_metric_kernels:
            line["metric_kernels"] = metric_kernel_lines(dev, peaks)
        if world == 1 and not args.no_cpu_baseline:
            sample_n = cpu_sample_size(wl)
            cpu_model = build_model(wl)
            rate, dt = time_cpu(wl, cpu_model, sample_n, 2, 1)
            line["cpu_baseline"] = {
                "value": rate, "unit": "sample*members/s", "cores": torch.get_num_threads(),
                "kind": "port",
                "sample": f"{sample_n} of {n} samples x {k} members, 1 warm-up + 2 timed calls, "
                          f"{dt:.2f} s per call, os.cpu_count={os.cpu_count()}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-metric-kernels", action="store_true",
                    help="skip the Wasserstein / KDE-JS kernel timings added to the default line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
