#!/usr/bin/env python
"""Golden vectors for KDEMLPModel's input-density score.

Runs the reference's OWN ``KDEMLPModel`` (``nnueehcs/models.py:191-222``), unmodified, from
``/root/reference``: ``fit_kde`` (sklearn ``KernelDensity(bandwidth='scott', rtol=1e-5)``) and
``forward(x, return_ue=True)``.  sklearn is the unpinned dependency that holds the arithmetic
(``pyproject.toml`` gives no version; the version used is stored in the file).  The random
``train_fit_prop`` subset is stored too, so the fixture does not depend on ``torch.randperm``.

    python tests/golden/make_golden_kde.py
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden.make_golden import mlp_arch, pack_state, ref_builder, ref_models  # noqa: E402


def make_case(out, tag, d, m, n, prop, seed, spread=1.0):
    import sklearn
    arch = mlp_arch(d, [16], 1, False)
    torch.manual_seed(seed)
    model = ref_builder.KDEModelBuilder(arch, {"bandwidth": "scott", "rtol": 0.1,
                                               "train_fit_prop": prop}).build()
    assert isinstance(model, ref_models.KDEMLPModel)
    g = torch.Generator().manual_seed(seed)
    train = torch.rand(m, d, generator=g) * spread
    # queries: in-distribution, shifted (OOD) and far away (density underflows to 0)
    x = torch.cat([torch.rand(n, d, generator=g) * spread,
                   torch.rand(n // 2, d, generator=g) * spread + 0.75 * spread,
                   torch.full((3, d), 40.0 * spread)])
    torch.manual_seed(seed + 1)
    model.fit_kde(train)
    model.eval()
    with torch.no_grad():
        pred, dens = model(x, return_ue=True)
    fitted = np.asarray(model.kde.tree_.data)          # the rows sklearn actually kept (float64)
    out[f"{tag}.fit"] = fitted.astype(np.float32)
    assert np.array_equal(out[f"{tag}.fit"].astype(np.float64), fitted)
    out[f"{tag}.x"] = x.numpy()
    out[f"{tag}.dens"] = dens.numpy()
    out[f"{tag}.pred"] = pred.numpy()
    out[f"{tag}.bandwidth"] = np.float64(model.kde.bandwidth_)
    out[f"{tag}.arch_yaml"] = yaml.safe_dump(arch)
    pack_state(f"{tag}.m0", model.model, out)
    out["sklearn_version"] = sklearn.__version__
    print(tag, "fit", fitted.shape, "h", model.kde.bandwidth_, "dens", dens[:3].tolist(),
          "min|dens|", float(dens.abs().min()))


if __name__ == "__main__":
    torch.set_num_threads(1)
    out = {"tags": np.array(["binomial5", "d2_half", "d12"])}
    make_case(out, "binomial5", 5, 3000, 200, 1.0, 0)       # binomial-options shaped inputs
    make_case(out, "d2_half", 2, 1500, 120, 0.5, 1, 3.0)    # train_fit_prop < 1, wider spread
    make_case(out, "d12", 12, 800, 64, 1.0, 2)              # runtime-width kernel (d > 8)
    np.savez_compressed(os.path.join(HERE, "kde_density.npz"), **out)
