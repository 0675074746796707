#!/usr/bin/env python
"""Golden vectors for PAGER, generated in the authoring container.

Runs the reference's OWN ``PAGERMLP`` class, unmodified, from ``/root/reference`` (built by its
own ``PAGERModelBuilder``), with the absent third-party ``deltauq.deltaUQ_MLP`` bound to this
repo's restatement (``oracle.shims``) -- so, like ``deltauq_small.npz``, the file pins the
reference's wrapper logic (``models.py:376-434``: Delta-UQ forward, swapped-role prediction matrix,
``max |P - Y|``, ``torch.maximum``) composed with the restated anchoring: PARITY UNPINNED for the
anchoring itself.  ``vectorize=False`` (the reference default): the loop branch of
``_anchored_predictions`` calls ``deltaUQ_MLP.forward`` with one "anchor" (the sample) at a time, so
the eval-mode anchor permutation inside ``deltaUQ_MLP`` cannot reorder anything.

    python tests/golden/make_golden_pager.py
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden.make_golden import (mlp_arch, pack_state, randomise_bn, ref_builder,  # noqa: E402
                                      ref_models)


def make_pager(name, d_in, widths, d_out, k, n, seed=0):
    arch = mlp_arch(d_in, widths, d_out, True)
    torch.manual_seed(13)
    model = ref_builder.PAGERModelBuilder(arch, {"estimator": "std", "num_anchors": k}).build()
    randomise_bn(model.net, 3100)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d_in, generator=g)
    anchors = torch.rand(k, d_in, generator=g)
    model.anchors = anchors
    model.eval()
    torch.manual_seed(5)
    with torch.no_grad():
        # targets near the predictions, so that the Delta-UQ std wins on some rows and the
        # conformal score on others (both sides of the torch.maximum are exercised)
        mu, std = ref_models.DeltaUQMLP.forward(model, x, True)
        anchors_y = mu.median() + (torch.rand(k, d_out, generator=g) - 0.5) * 2.0 * std.median()
        model.anchors_Y = anchors_y
        pred, score = model(x, return_ue=True)
        conformal = model._score_samples(x, model.anchors, model.anchors_Y)
        pred_only = model(x)
    out = {"arch_yaml": yaml.safe_dump(arch), "k": k, "x": x.numpy(), "anchors": anchors.numpy(),
           "anchors_y": anchors_y.numpy(), "pred": pred.numpy(), "score": score.numpy(),
           "conformal": conformal.numpy(), "pred_only": pred_only.numpy(), "std": std.numpy()}
    pack_state("m0", model.net, out)
    np.savez_compressed(os.path.join(HERE, name), **out)
    frac = float((conformal > std).float().mean())
    print(name, "pred", float(pred.abs().mean()), "score", float(score.mean()),
          "conformal wins on", frac, "of rows")


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_pager("pager_small.npz", 5, [64, 64], 1, k=6, n=150)
