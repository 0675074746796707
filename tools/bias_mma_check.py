"""A/B of the bias-in-the-MMA variant of uq_mlp_tc2_kernel (UQ_TC_BIAS_MMA=1) against the
epilogue-bias variant (=0) and the float64-accumulating oracle, on ensembles of every pair-kernel
width with LARGE biases (a dropped or mis-placed bias piece is then an O(1) error)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnueehcs_b200 import ops  # noqa: E402
from oracle import uq_oracle  # noqa: E402

DEV = torch.device("cuda:0")


def nets_of(width, n_hidden, k, d_in, bias_scale):
    nets = []
    for i in range(k):
        torch.manual_seed(7 + i)
        layers, fan = [], d_in
        for _ in range(n_hidden):
            layers += [torch.nn.Linear(fan, width), torch.nn.ReLU()]
            fan = width
        layers += [torch.nn.Linear(fan, 1)]
        net = torch.nn.Sequential(*layers).eval()
        with torch.no_grad():
            for m in net:
                if isinstance(m, torch.nn.Linear):
                    m.bias.mul_(bias_scale).add_(0.37 * bias_scale * torch.randn_like(m.bias))
        nets.append(net)
    return nets


def run(packed, x, k, flag):
    os.environ["UQ_TC_BIAS_MMA"] = flag
    mean, std = packed.forward(x, "ensemble", total_members=k, precision="bf16")
    torch.cuda.synchronize()
    return mean.double().cpu(), std.double().cpu()


ok = True
for width, n_hidden, k, n in [(192, 2, 3, 257), (256, 3, 4, 1030), (320, 2, 2, 383), (384, 4, 3, 999),
                              (448, 2, 2, 129), (512, 3, 4, 5000), (512, 1, 2, 300), (256, 6, 5, 70000)]:
    for bias_scale in (1.0, 20.0):
        nets = nets_of(width, n_hidden, k, 5, bias_scale)
        x = torch.rand(n, 5, generator=torch.Generator().manual_seed(n))
        packed = ops.PackedModel(nets, DEV)
        ref_mean, ref_std = uq_oracle.ensemble_forward(nets, x)
        ref_mean, ref_std = torch.as_tensor(ref_mean).double(), torch.as_tensor(ref_std).double()
        scale = float(ref_mean.abs().max() + ref_std.abs().max())
        m0, s0 = run(packed, x.to(DEV), k, "0")
        m1, s1 = run(packed, x.to(DEV), k, "1")
        e0 = max(float((m0 - ref_mean).abs().max()), float((s0 - ref_std).abs().max())) / scale
        e1 = max(float((m1 - ref_mean).abs().max()), float((s1 - ref_std).abs().max())) / scale
        ab = max(float((m1 - m0).abs().max()), float((s1 - s0).abs().max())) / scale
        good = e1 <= max(2.0 * e0, 3e-3) and e1 < 1e-2
        ok &= good
        print(f"{'OK  ' if good else 'FAIL'} H={width} L={n_hidden} k={k} n={n} bias x{bias_scale}: "
              f"err vs oracle epilogue-bias {e0:.3e}, bias-in-MMA {e1:.3e}, A/B diff {ab:.3e} (of scale)")

# narrow nets (mlp_tc4.cu) as ensembles
for width, n_hidden, k, n in [(64, 2, 3, 1), (64, 3, 2, 1025), (128, 2, 4, 2049), (128, 6, 5, 30000)]:
    for bias_scale in (1.0, 20.0):
        nets = nets_of(width, n_hidden, k, 5, bias_scale)
        x = torch.rand(n, 5, generator=torch.Generator().manual_seed(n))
        packed = ops.PackedModel(nets, DEV)
        ref_mean, ref_std = uq_oracle.ensemble_forward(nets, x)
        ref_mean, ref_std = torch.as_tensor(ref_mean).double(), torch.as_tensor(ref_std).double()
        scale = float(ref_mean.abs().max() + ref_std.abs().max())
        m0, s0 = run(packed, x.to(DEV), k, "0")
        m1, s1 = run(packed, x.to(DEV), k, "1")
        e0 = max(float((m0 - ref_mean).abs().max()), float((s0 - ref_std).abs().max())) / scale
        e1 = max(float((m1 - ref_mean).abs().max()), float((s1 - ref_std).abs().max())) / scale
        ab = max(float((m1 - m0).abs().max()), float((s1 - s0).abs().max())) / scale
        good = e1 <= max(2.0 * e0, 3e-3) and e1 < 1e-2
        ok &= good
        print(f"{'OK  ' if good else 'FAIL'} narrow H={width} L={n_hidden} k={k} n={n} bias x{bias_scale}: "
              f"err vs oracle epilogue-bias {e0:.3e}, bias-in-MMA {e1:.3e}, A/B diff {ab:.3e} (of scale)")

# anchored modes: per-anchor layer-0 bias stages built per call (narrow and pair kernels)
for width, n_hidden, k, n in [(128, 6, 32, 20011), (64, 2, 5, 700), (256, 3, 7, 1500), (512, 2, 4, 900)]:
    for bias_scale in (1.0, 20.0):
        net = nets_of(width, n_hidden, 1, 10, bias_scale)[0]
        x = torch.rand(n, 5, generator=torch.Generator().manual_seed(n))
        anchors = torch.rand(k, 5, generator=torch.Generator().manual_seed(2)) * 3.0
        packed = ops.PackedModel([net], DEV)
        ref_mean, ref_std = uq_oracle.delta_uq_forward(net, x, anchors, k)
        ref_mean, ref_std = torch.as_tensor(ref_mean).double(), torch.as_tensor(ref_std).double()
        scale = float(ref_mean.abs().max() + ref_std.abs().max())
        outs = {}
        for flag in ("0", "1"):
            os.environ["UQ_TC_BIAS_MMA"] = flag
            mean, std = packed.forward(x.to(DEV), "delta_uq", total_members=k, precision="bf16",
                                       anchors=anchors.to(DEV))
            torch.cuda.synchronize()
            outs[flag] = (mean.double().cpu(), std.double().cpu())
        e = {f: max(float((outs[f][0] - ref_mean).abs().max()),
                    float((outs[f][1] - ref_std).abs().max())) / scale for f in outs}
        ab = max(float((outs["1"][0] - outs["0"][0]).abs().max()),
                 float((outs["1"][1] - outs["0"][1]).abs().max())) / scale
        good = e["1"] <= max(2.0 * e["0"], 3e-3) and e["1"] < 1e-2
        ok &= good
        print(f"{'OK  ' if good else 'FAIL'} delta-uq H={width} L={n_hidden} anchors={k} n={n} bias x{bias_scale}: "
              f"err vs oracle epilogue-bias {e['0']:.3e}, bias-in-MMA {e['1']:.3e}, A/B diff {ab:.3e}")
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
