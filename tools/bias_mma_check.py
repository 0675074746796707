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
                              (448, 2, 2, 129), (512, 3, 4, 5000), (512, 1, 2, 300), (256, 6, 5, 70000),
                              (1024, 2, 3, 300), (768, 3, 2, 193), (1024, 7, 2, 20000)]:
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
for width, n_hidden, k, n in [(128, 6, 32, 20011), (64, 2, 5, 700), (256, 3, 7, 1500), (512, 2, 4, 900),
                              (1024, 3, 5, 333)]:
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

# live dropout: native Philox masks exported and replayed through the oracle, both variants
import numpy as np  # noqa: E402


def mc_net(width, n_hidden, p, bias_scale, final_drop=False):
    torch.manual_seed(3)
    layers, fan = [], 5
    for i in range(n_hidden):
        layers += [torch.nn.Linear(fan, width), torch.nn.ReLU()]
        if i > 0 or final_drop:
            layers += [torch.nn.Dropout(p)]
        fan = width
    if not final_drop and n_hidden > 1:
        layers.pop()          # the reference's builder puts no dropout before the final Linear
    layers += [torch.nn.Linear(fan, 1)]
    net = torch.nn.Sequential(*layers).eval()
    with torch.no_grad():
        for m in net:
            if isinstance(m, torch.nn.Linear):
                m.bias.mul_(bias_scale).add_(0.37 * bias_scale * torch.randn_like(m.bias))
    return net


def masks_from_flat(flat, n, widths, passes):
    out, off = [], 0
    flat = flat.cpu().numpy()
    for w in widths:
        cnt = passes * n * w
        out.append(torch.from_numpy(flat[off:off + cnt].reshape(passes, n, w).astype(np.float32)))
        off += cnt
    return [[m[k] for m in out] for k in range(passes)]


for width, n_hidden, passes, n, final_drop in [(128, 6, 9, 700, False), (64, 3, 5, 300, True),
                                               (256, 4, 6, 500, False), (512, 3, 5, 400, True),
                                               (320, 3, 4, 260, False), (1024, 4, 5, 300, False),
                                               (768, 3, 4, 200, True)]:
    pdrop, seed = 0.2, 1234 + width
    net = mc_net(width, n_hidden, pdrop, 5.0, final_drop)
    x = torch.rand(n, 5, generator=torch.Generator().manual_seed(n))
    packed = ops.PackedModel([net], DEV)
    flat = ops.philox_keep_masks(n, packed.dropout_widths, passes, pdrop, seed, 0, DEV)
    try:
        from tests.util import injected_to_masks
        masks = injected_to_masks(flat.cpu(), n, packed.dropout_widths, passes)
    except Exception:
        masks = masks_from_flat(flat, n, packed.dropout_widths, passes)
    ref_mean, ref_std = uq_oracle.mc_dropout_forward(net, x, passes, pdrop, masks=masks)
    ref_mean, ref_std = torch.as_tensor(ref_mean).double(), torch.as_tensor(ref_std).double()
    scale = float(ref_mean.abs().max() + ref_std.abs().max())
    outs = {}
    for flag in ("0", "1"):
        os.environ["UQ_TC_BIAS_MMA"] = flag
        mean, std = packed.forward(x.to(DEV), "mc_dropout", total_members=passes, precision="bf16",
                                   dropout_p=pdrop, seed=seed)
        torch.cuda.synchronize()
        outs[flag] = (mean.double().cpu(), std.double().cpu())
    e = {f: max(float((outs[f][0] - ref_mean).abs().max()),
                float((outs[f][1] - ref_std).abs().max())) / scale for f in outs}
    ab = max(float((outs["1"][0] - outs["0"][0]).abs().max()),
             float((outs["1"][1] - outs["0"][1]).abs().max())) / scale
    good = e["1"] <= max(2.0 * e["0"], 3e-3) and e["1"] < 1e-2
    ok &= good
    print(f"{'OK  ' if good else 'FAIL'} mc-dropout H={width} L={n_hidden} passes={passes} n={n} "
          f"final_drop={final_drop}: err vs oracle epilogue-bias {e['0']:.3e}, bias-in-MMA {e['1']:.3e}, "
          f"A/B diff {ab:.3e}")
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
