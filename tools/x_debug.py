"""Bring-up check of the fp32-parity split kernel (csrc/mlp_tcx.cu): for a list of shapes, run
precision='fp32' (tensor-core split), 'fp32_ffma' (CUDA cores) and 'bf16' against the CPU oracle and
print the worst errors in units of the 1e-5 tolerance.  GPU only; not part of the test-suite."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnueehcs_b200 import ops  # noqa: E402
from nnueehcs_b200.model_builder import build_network  # noqa: E402
from oracle import uq_oracle  # noqa: E402

DEV = torch.device("cuda:0")


def arch(d_in, width, n_hidden, d_out, bn=True, drop=None):
    a, prev = [], d_in
    for i in range(n_hidden):
        if drop is not None and i > 0:
            a.append({"Dropout": {"args": [drop]}})
        a.append({"Linear": {"args": [prev, width]}})
        if bn:
            a.append({"BatchNorm1d": {"args": [width]}})
        a.append({"ReLU": {"inplace": True}})
        prev = width
    if drop is not None:
        a.append({"Dropout": {"args": [drop]}})
    a.append({"Linear": {"args": [prev, d_out]}})
    return a


def rand_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)


def excess(got, ref, scale):
    """max over elements of |got - ref| / (1e-5 |ref| + 1e-5 max|scale|): <= 1 passes"""
    got, ref = got.double().cpu(), torch.as_tensor(ref).double()
    tol = 1e-5 * ref.abs() + 1e-5 * float(torch.as_tensor(scale).abs().max())
    return float(((got - ref).abs() / tol).max())


def main():
    shapes = [(64, 2, 3, 100, 1), (128, 6, 3, 1000, 1), (192, 2, 2, 257, 1), (256, 3, 4, 130, 1),
              (320, 2, 2, 383, 1), (384, 2, 2, 200, 3), (448, 2, 2, 129, 1), (512, 3, 4, 4097, 1),
              (512, 3, 16, 20000, 1), (128, 6, 32, 300, 1)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in s.split(",")) for s in sys.argv[1:]]
    for width, n_hidden, k, n, d_out in shapes:
        nets = []
        for i in range(k):
            torch.manual_seed(42 + i)
            net = build_network(arch(5, width, n_hidden, d_out)).eval()
            rand_bn(net, 1 + i)
            nets.append(net)
        x = torch.rand(n, 5, generator=torch.Generator().manual_seed(0))
        packed = ops.PackedModel(nets, DEV)
        t0 = time.time()
        ref_mean, ref_std = uq_oracle.ensemble_forward(nets, x)
        t_cpu = time.time() - t0
        line = f"H={width} L={n_hidden} K={k} n={n} dout={d_out} split={packed.fp32_on_tensor_cores}"
        for prec in ("fp32", "fp32_ffma", "bf16"):
            try:
                mean, std = packed.forward(x.to(DEV), "ensemble", total_members=k, precision=prec)
                torch.cuda.synchronize()
                sc = float(ref_mean.abs().max())
                d = (mean.double().cpu() - ref_mean.double()) / sc
                line += (f" | {prec}: mean {excess(mean, ref_mean, ref_mean):.3f} "
                         f"std {excess(std, ref_std, ref_mean):.3f} "
                         f"[bias {float(d.mean()):+.2e} rms {float(d.pow(2).mean().sqrt()):.2e}]")
            except Exception as e:  # noqa: BLE001
                line += f" | {prec}: {type(e).__name__} {e}"
        print(line + f" | cpu {t_cpu:.1f}s", flush=True)
    # MC dropout with injected Philox masks on the split path (6 x 128, p = 0.2)
    torch.manual_seed(42)
    net = build_network(arch(5, 128, 6, 1, drop=0.2)).eval()   # incl. a Dropout before the last Linear
    rand_bn(net, 1)
    packed = ops.PackedModel([net], DEV)
    n, passes, seed = 300, 12, 7
    x = torch.rand(n, 5, generator=torch.Generator().manual_seed(3))
    flat = ops.philox_keep_masks(n, packed.dropout_widths, passes, 0.2, seed, 0, DEV)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from tests.util import injected_to_masks
    masks = injected_to_masks(flat.cpu(), n, packed.dropout_widths, passes)
    ref_mean, ref_std = uq_oracle.mc_dropout_forward(net, x, passes, 0.2, masks=masks)
    for prec in ("fp32", "fp32_ffma"):
        mean, std = packed.forward(x.to(DEV), "mc_dropout", total_members=passes, precision=prec,
                                   dropout_p=0.2, seed=seed)
        print(f"mc_dropout philox {prec}: mean {excess(mean, ref_mean, ref_mean):.3f} "
              f"std {excess(std, ref_std, ref_mean):.3f}", flush=True)
    # Delta-UQ on the split path
    torch.manual_seed(42)
    a = arch(10, 128, 6, 1)
    net = build_network(a).eval()
    rand_bn(net, 1)
    packed = ops.PackedModel([net], DEV)
    x = torch.rand(500, 5, generator=torch.Generator().manual_seed(4))
    anchors = torch.rand(8, 5, generator=torch.Generator().manual_seed(5))
    ref_mean, ref_std = uq_oracle.delta_uq_forward(net, x, anchors, 8)
    for prec in ("fp32", "fp32_ffma"):
        mean, std = packed.forward(x.to(DEV), "delta_uq", total_members=8, precision=prec,
                                   anchors=anchors.to(DEV))
        print(f"delta_uq {prec}: mean {excess(mean, ref_mean, ref_mean):.3f} "
              f"std {excess(std, ref_std, ref_mean):.3f}", flush=True)


if __name__ == "__main__":
    main()
