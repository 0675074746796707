"""Host-side mirror of ``nnueehcs/classification.py``: percentile-threshold ID/OOD classifiers.

Same class names, constructor checks, ``evaluate`` / ``_evaluate_scores`` contract and result keys
as the reference (classification.py:5-196); the threshold (``torch.quantile`` of the ID scores)
and the four counts come from one call into ``uq_score_metrics`` (device sorts + binary searches)
instead of ``torch.quantile`` plus four boolean reductions with ``.item()`` syncs.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class _IdOodClassifier:
    def evaluate(self, model: nn.Module, id_data: tuple, ood_data: tuple) -> dict:
        id_ipt, _ = id_data
        ood_ipt, _ = ood_data
        model.eval()
        with torch.no_grad():
            id_preds, id_scores = model(id_ipt, return_ue=True)
            ood_preds, ood_scores = model(ood_ipt, return_ue=True)
        metrics = self._evaluate_scores(id_scores, ood_scores)
        metrics.update({"id_preds": id_preds, "ood_preds": ood_preds, "id_scores": id_scores,
                        "ood_scores": ood_scores})
        return metrics


def _cuda(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("nnueehcs_b200 metrics run on CUDA only (no CPU fallback)")
        t = t.cuda()
    return t


class PercentileBasedIdOodClassifier(_IdOodClassifier):
    """classification.py:28-151: a sample is flagged OOD when its score exceeds the
    ``percentile`` quantile of the ID scores."""

    def __init__(self, percentile: float):
        if not 0 <= percentile <= 1:
            raise ValueError(f"Percentile must be between 0 and 1, got {percentile}")
        super().__init__()
        self.percentile = percentile

    def _evaluate_scores(self, id_scores: torch.Tensor, ood_scores: torch.Tensor) -> dict:
        r = ops.score_metrics(_cuda(id_scores), _cuda(ood_scores),
                              classifier_percentile=self.percentile)
        return {k: r[k] for k in ("sensitivity", "specificity", "fpr", "fnr")}

    @classmethod
    def get_objectives(cls):
        return [{'name': 'sensitivity', 'type': 'maximize'}]

    @classmethod
    def get_metrics(cls):
        return ['sensitivity']


class ReversedPercentileBasedIdOodClassifier(PercentileBasedIdOodClassifier):
    """classification.py:154-196: lower scores indicate OOD; threshold at the ``1 - percentile``
    quantile, positives are the OOD scores at or below it."""

    def _evaluate_scores(self, id_scores: torch.Tensor, ood_scores: torch.Tensor) -> dict:
        r = ops.score_metrics(_cuda(id_scores), _cuda(ood_scores),
                              classifier_percentile=1 - self.percentile)
        # reference: FP = #(id <= t), FN = #(ood > t), TP = #(ood <= t), TN = #(id > t)
        return {"sensitivity": r["fnr"], "specificity": r["fpr"], "fpr": r["specificity"],
                "fnr": r["sensitivity"]}
