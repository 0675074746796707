#!/usr/bin/env python
"""Generate the golden fixtures that pin the oracle (and, through it, the CUDA path).

Runs the reference's OWN classes, unmodified, from ``/root/reference`` (via ``oracle.shims``)
and scipy, and writes small ``.npz`` files next to this script.  The reference tree cannot
travel to the GPU box, the fixtures can.  Re-run with::

    python tests/golden/make_golden.py

What each fixture holds (all float32 unless noted):

``ensemble_*.npz``   arch (YAML text), K members' ``state_dict`` built by the reference
                     ``EnsembleModelBuilder`` (seeds 42+i, model_builder.py:229) with BatchNorm
                     running stats randomised, inputs ``x``, per-member outputs, and the
                     reference ``EnsembleModel.forward(x, return_ue=True)`` mean/std.
``mcdropout_*.npz``  same for ``MCDropoutModelBuilder``: the dropout keep-masks harvested from the
                     reference's own run (forward hooks), its mean/std, and the dropout-off
                     mean/std (``nn.Module.eval`` semantics: P identical passes).
``deltauq_*.npz``    reference ``DeltaUQMLP`` wrapper (chunk/concat logic, models.py:313-341)
                     composed with this repo's RESTATED ``deltaUQ_MLP`` -- parity unpinned.
``metrics_*.npz``    scipy ``wasserstein_distance`` and the reference's
                     ``JensenShannonEvaluation`` on seeded Gamma samples (float64 results).
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import shims  # noqa: E402
from oracle import uq_oracle  # noqa: E402

ref_models, ref_builder, ref_eval = shims.import_reference()


def mlp_arch(d_in, widths, d_out, bn=True):
    arch = []
    prev = d_in
    for w in widths:
        arch.append({"Linear": {"args": [prev, w]}})
        if bn:
            arch.append({"BatchNorm1d": {"args": [w]}})
        arch.append({"ReLU": {"inplace": True}})
        prev = w
    arch.append({"Linear": {"args": [prev, d_out]}})
    return arch


def randomise_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))


def pack_state(prefix, net, out):
    for k, v in net.state_dict().items():
        out[f"{prefix}.{k}"] = v.detach().cpu().numpy()


def make_ensemble(name, d_in, widths, d_out, k, n, bn=True, seed=0):
    arch = mlp_arch(d_in, widths, d_out, bn)
    model = ref_builder.EnsembleModelBuilder(arch, {"num_models": k}).build()
    for i, net in enumerate(model.models):
        randomise_bn(net, 1000 + i)
    model.eval()
    x = torch.rand(n, d_in, generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        mean, std = model(x, return_ue=True)
        members = torch.stack([net(x) for net in model.models])
    out = {"arch_yaml": yaml.safe_dump(arch), "k": k, "x": x.numpy(),
           "mean": mean.numpy(), "std": std.numpy(), "members": members.numpy()}
    for i, net in enumerate(model.models):
        pack_state(f"m{i}", net, out)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "mean", float(mean.abs().mean()), "std", float(std.mean()))


def make_mcdropout(name, d_in, widths, d_out, p, passes, n, bn=True, seed=0):
    arch = mlp_arch(d_in, widths, d_out, bn)
    torch.manual_seed(7)
    model = ref_builder.MCDropoutModelBuilder(
        arch, {"num_samples": passes, "dropout_percent": p}).build()
    randomise_bn(model.model, 2000)
    x = torch.rand(n, d_in, generator=torch.Generator().manual_seed(seed))
    # reference semantics: MCDropoutModel.eval() keeps Dropout live (models.py:165-169)
    model.eval()
    torch.manual_seed(123)
    masks, (mean, std) = uq_oracle.harvest_masks(model, x, passes)
    # dropout-off: plain nn.Module.eval() semantics -> P identical passes
    torch.nn.Module.eval(model)
    with torch.no_grad():
        mean_off, std_off = model(x, return_ue=True)
    full_arch = [{type(m).__name__: None} for m in model.model]
    nl = len(masks[0])
    out = {"arch_yaml": yaml.safe_dump(arch), "p": p, "passes": passes, "x": x.numpy(),
           "mean": mean.numpy(), "std": std.numpy(),
           "mean_off": mean_off.numpy(), "std_off": std_off.numpy(),
           "n_dropout_layers": nl,
           "module_names": np.array([list(d.keys())[0] for d in full_arch])}
    for l in range(nl):
        out[f"mask_l{l}"] = np.packbits(
            torch.stack([masks[s][l] for s in range(passes)]).numpy(), axis=-1)
        out[f"mask_l{l}_width"] = masks[0][l].shape[-1]
    pack_state("m0", model.model, out)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "mean", float(mean.abs().mean()), "std", float(std.mean()),
          "std_off", float(std_off.abs().max()))


def make_deltauq(name, d_in, widths, d_out, k, n, chunk, bn=True, seed=0):
    arch = mlp_arch(d_in, widths, d_out, bn)
    torch.manual_seed(11)
    model = ref_builder.DeltaUQMLPModelBuilder(
        arch, {"estimator": "std", "num_anchors": k, "anchored_batch_size": chunk}).build()
    randomise_bn(model.net, 3000)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d_in, generator=g)
    anchors = torch.rand(k, d_in, generator=g)
    model.anchors = anchors
    model.eval()
    torch.manual_seed(5)
    with torch.no_grad():
        mean, std = model(x, return_ue=True)
    out = {"arch_yaml": yaml.safe_dump(arch), "k": k, "chunk": chunk, "x": x.numpy(),
           "anchors": anchors.numpy(), "mean": mean.numpy(), "std": std.numpy()}
    pack_state("m0", model.net, out)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "mean", float(mean.abs().mean()), "std", float(std.mean()))


def make_metrics(name, n_id, n_ood, seed=0):
    from scipy.stats import wasserstein_distance
    rng = np.random.default_rng(seed)
    id_s = rng.gamma(2.0, 0.05, n_id).astype(np.float32)
    ood_s = rng.gamma(3.0, 0.08, n_ood).astype(np.float32)
    w = wasserstein_distance(id_s, ood_s)
    w_same = wasserstein_distance(id_s, id_s)
    js = ref_eval.JensenShannonEvaluation()
    jsd = js.pdf_jsd(id_s, ood_s)
    # through the metric classes themselves (UncertaintyEstimate hop included)
    w_cls = ref_eval.WassersteinEvaluation()._evaluate_uncertainties(
        ref_eval.UncertaintyEstimate(id_s[:, None]), ref_eval.UncertaintyEstimate(ood_s[:, None]))
    j_cls = js._evaluate_uncertainties(
        ref_eval.UncertaintyEstimate(id_s[:, None]), ref_eval.UncertaintyEstimate(ood_s[:, None]))
    assert w_cls["wasserstein_distance"] == w and j_cls["jensen_shannon_distance"] == jsd
    np.savez_compressed(os.path.join(HERE, name), id=id_s, ood=ood_s,
                        wasserstein=np.float64(w), wasserstein_same=np.float64(w_same),
                        jsd=np.float64(jsd))
    print(name, "W", w, "W_same", w_same, "JSD", jsd)


if __name__ == "__main__":
    torch.set_num_threads(1)
    # small, odd widths (the reference's own test nets use 25-wide layers, out=5)
    make_ensemble("ensemble_small.npz", 16, [25, 25, 25], 5, k=4, n=96, bn=False)
    make_ensemble("ensemble_bn.npz", 5, [64, 64], 1, k=5, n=200, bn=True)
    # canonical binomial-options net (examples/binomial_options/config.yaml:16-54)
    make_ensemble("ensemble_binomial.npz", 5, [128] * 6, 1, k=3, n=160, bn=True)
    make_mcdropout("mcdropout_small.npz", 16, [25, 25, 25], 5, p=0.2, passes=6, n=96, bn=False)
    make_mcdropout("mcdropout_binomial.npz", 5, [128] * 6, 1, p=0.2, passes=8, n=160, bn=True)
    make_deltauq("deltauq_small.npz", 5, [64, 64], 1, k=6, n=150, chunk=64, bn=True)
    make_metrics("metrics_small.npz", 3000, 2500, seed=0)
    make_metrics("metrics_medium.npz", 20000, 20000, seed=1)
