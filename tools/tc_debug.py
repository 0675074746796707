#!/usr/bin/env python
"""Progressive bring-up check of the bf16 tcgen05 kernel against the CPU oracle (GPU box only).

Runs from the simplest shape (one MMA layer, one tile) to the benchmark shape and prints, for the
first failing case, where in the [row, hidden-unit] plane the errors sit -- which tells a
descriptor / swizzle / TMEM-lane mistake from a protocol one."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnueehcs_b200 import ops  # noqa: E402
from oracle import uq_oracle  # noqa: E402

DEV = torch.device("cuda:0")


def make_net(d_in, widths, d_out, seed, bn=False):
    torch.manual_seed(seed)
    layers, prev = [], d_in
    for w in widths:
        layers.append(nn.Linear(prev, w))
        if bn:
            b = nn.BatchNorm1d(w)
            b.running_mean.normal_(0, 0.1)
            b.running_var.uniform_(0.5, 1.5)
            layers.append(b)
        layers.append(nn.ReLU())
        prev = w
    layers.append(nn.Linear(prev, d_out))
    return nn.Sequential(*layers).eval()


def hidden_probe(widths, n=128):
    """Read every last-hidden unit through one-hot output weights (d_out = 8 at a time)."""
    net = make_net(5, widths, 8, 0)
    x = torch.rand(n, 5, generator=torch.Generator().manual_seed(0))
    H = widths[-1]
    got = torch.zeros(n, H)
    ref = torch.zeros(n, H)
    for j0 in range(0, H, 8):
        with torch.no_grad():
            net[-1].weight.zero_()
            net[-1].bias.zero_()
            for o in range(8):
                net[-1].weight[o, j0 + o] = 1.0
        packed = ops.PackedModel([net, net], DEV)
        mean, _ = packed.forward(x.to(DEV), "ensemble", total_members=2, precision="bf16")
        got[:, j0:j0 + 8] = mean.cpu()
        ref[:, j0:j0 + 8] = uq_oracle.ensemble_forward([net, net], x)[0]
        packed.close()
    err = (got - ref).abs()
    scale = float(ref.abs().max())
    print(f"  hidden probe {widths}: max err {float(err.max()):.3e} (scale {scale:.3e})")
    if float(err.max()) > 3e-2 * scale:
        bad = err > 3e-2 * scale
        print("  bad fraction", float(bad.float().mean()))
        print("  bad by row%8  :", [round(float(bad[r::8].float().mean()), 3) for r in range(8)])
        print("  bad by row//32:", [round(float(bad[32 * r:32 * r + 32].float().mean()), 3) for r in range(n // 32)])
        print("  bad by col%8  :", [round(float(bad[:, c::8].float().mean()), 3) for c in range(8)])
        print("  bad by col//8 :", [round(float(bad[:, 8 * c:8 * c + 8].float().mean()), 3) for c in range(H // 8)])
        print("  got[0,:8]", got[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        print("  got[1,:8]", got[1, :8].tolist())
        print("  ref[1,:8]", ref[1, :8].tolist())
        return False
    return True


def case(name, widths, k, n, bn=True, d_out=1, tol=3e-2):
    nets = [make_net(5, widths, d_out, 10 + i, bn) for i in range(k)]
    x = torch.rand(n, 5, generator=torch.Generator().manual_seed(1))
    packed = ops.PackedModel(nets, DEV)
    mean, std = packed.forward(x.to(DEV), "ensemble", total_members=k, precision="bf16")
    m32, s32 = packed.forward(x.to(DEV), "ensemble", total_members=k, precision="fp32")
    rm, rs = uq_oracle.ensemble_forward(nets, x)
    sc = float(rm.abs().max())
    e16 = float((mean.cpu() - rm).abs().max())
    e32 = float((m32.cpu() - rm).abs().max())
    es16 = float((std.cpu() - rs).abs().max()) if k > 1 else 0.0
    ok = e16 <= tol * sc
    print(f"{'OK  ' if ok else 'FAIL'} {name}: bf16 mean err {e16:.3e}, std err {es16:.3e}, "
          f"fp32 mean err {e32:.3e} (scale {sc:.3e})")
    packed.close()
    return ok


def main():
    torch.set_num_threads(8)
    ok = True
    ok &= hidden_probe([64])
    ok &= hidden_probe([64, 64])
    ok &= hidden_probe([128, 128])
    ok &= hidden_probe([512, 512])
    ok &= case("1x[64] n=128", [64], 2, 128)
    ok &= case("2x[64,64] n=128", [64, 64], 2, 128)
    ok &= case("2x[128]*3 n=128", [128] * 3, 2, 128)
    ok &= case("3x[256]*2 n=300", [256] * 2, 3, 300)
    ok &= case("4x[512]*3 n=1000", [512] * 3, 4, 1000)
    ok &= case("16x[512]*3 n=40000", [512] * 3, 16, 40000)
    ok &= case("2x[192]*2 n=200 d_out=3", [192] * 2, 2, 200, d_out=3)
    ok &= case("2x[384]*2 n=200", [384] * 2, 2, 200)
    if os.environ.get("UQ_TC_DEBUG_WIDE", "1") == "1":
        ok &= hidden_probe([1024, 1024], n=64)
        ok &= hidden_probe([1024, 1024], n=128)
        ok &= case("2x[1024]*1 n=64", [1024], 2, 64)
        ok &= case("2x[1024]*2 n=64", [1024] * 2, 2, 64)
        ok &= case("3x[1024]*3 n=1000", [1024] * 3, 3, 1000)
        ok &= case("2x[768]*3 n=300 d_out=2", [768] * 3, 2, 300, d_out=2)
        ok &= case("4x[1024]*7 n=20000", [1024] * 7, 4, 20000)
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
