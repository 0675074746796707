"""nnueehcs_b200 -- B200-native (sm_100a) uncertainty-estimation inference for NNUEEHCS models.

Drop-in for one hot path of cjlauer16/NNUEEHCS: the repeated forward passes behind the
MC-dropout / deep-ensemble / Delta-UQ wrappers, their per-sample mean/std reduction, and the
Wasserstein and KDE-Jensen-Shannon metrics over the resulting scores.  The Python modules mirror
the reference's ``model_builder`` / ``models`` / ``evaluation`` interface; the arithmetic lives in
``libnnueehcs_b200.so`` (hand-written CUDA behind the C ABI of ``include/nnueehcs_b200.h``).
"""
from . import evaluation, model_builder, models  # noqa: F401

__all__ = ["evaluation", "model_builder", "models"]
