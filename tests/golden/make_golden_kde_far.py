#!/usr/bin/env python
"""Golden vectors for KDEMLPModel's density score on FAR-OOD queries.

Same recipe as make_golden_kde.py -- the reference's OWN ``KDEMLPModel`` (``nnueehcs/models.py:
191-222``) from ``/root/reference``, sklearn ``KernelDensity`` underneath -- but the queries walk
away from the fitted cloud in steps, so their densities cover 1e-10 ... 1e-300 and beyond: the
range in which a float32 kernel sum has flushed to zero while the reference (log-space
``score_samples``, float64 ``exp``) still tells the scores apart.

    python tests/golden/make_golden_kde_far.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden.make_golden import mlp_arch, ref_builder, ref_models  # noqa: E402


def make_case(out, tag, d, m, seed, bandwidth="scott"):
    import sklearn
    arch = mlp_arch(d, [16], 1, False)
    torch.manual_seed(seed)
    model = ref_builder.KDEModelBuilder(arch, {"bandwidth": bandwidth, "rtol": 0.1,
                                               "train_fit_prop": 1.0}).build()
    assert isinstance(model, ref_models.KDEMLPModel)
    g = torch.Generator().manual_seed(seed)
    train = torch.rand(m, d, generator=g)
    # a ray leaving the unit cube: distance 0 ... 14 bandwidth-ish units from its corner
    steps = torch.linspace(0.0, 1.0, 97).unsqueeze(1)
    direction = torch.rand(1, d, generator=g) + 0.5
    x = torch.cat([torch.rand(16, d, generator=g),
                   1.0 + steps * direction * 9.0,
                   -steps * direction.flip(1) * 9.0])
    torch.manual_seed(seed + 1)
    model.fit_kde(train)
    model.eval()
    with torch.no_grad():
        _, dens = model(x, return_ue=True)
    fitted = np.asarray(model.kde.tree_.data)
    out[f"{tag}.fit"] = fitted.astype(np.float32)
    assert np.array_equal(out[f"{tag}.fit"].astype(np.float64), fitted)
    out[f"{tag}.x"] = x.numpy()
    out[f"{tag}.dens"] = dens.numpy()
    out[f"{tag}.bandwidth"] = np.float64(model.kde.bandwidth_)
    out["sklearn_version"] = sklearn.__version__
    a = np.abs(dens.numpy())
    print(tag, "fit", fitted.shape, "h", model.kde.bandwidth_, "|dens| range", a.max(),
          a[a > 0].min(), "zeros", int((a == 0).sum()),
          "below 1e-38:", int(((a < 1e-38) & (a > 0)).sum()))


if __name__ == "__main__":
    torch.set_num_threads(1)
    out = {"tags": np.array(["far5", "far2", "silverman3"])}
    make_case(out, "far5", 5, 2000, 10)
    make_case(out, "far2", 2, 700, 11)
    # bandwidth='silverman' (examples/bo_driven/config_kde.yaml:385-390 offers both rules)
    make_case(out, "silverman3", 3, 900, 12, "silverman")
    np.savez_compressed(os.path.join(HERE, "kde_density_far.npz"), **out)
