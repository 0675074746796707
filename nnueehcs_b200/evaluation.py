"""Host-side mirror of the hot-path part of ``nnueehcs/evaluation.py``.

Same class names, ``evaluate`` / ``get_objectives`` / ``get_metrics`` / ``get_name`` contract and
result keys as the reference; what changes is underneath:

* the scores returned by ``model(x, return_ue=True)`` stay on the GPU -- the reference's
  ``UncertaintyEstimate._to_numpy`` hop (``evaluation.py:88``: ``.detach().cpu().numpy()``) is
  gone; ``.data`` still hands out numpy on demand for callers that want it;
* ``WassersteinEvaluation._evaluate_uncertainties`` (``evaluation.py:175-188``) calls the CUDA
  sort + merged-CDF kernel instead of ``scipy.stats.wasserstein_distance``;
* ``JensenShannonEvaluation.pdf_jsd`` (``evaluation.py:268-276``) calls the CUDA KDE-on-grid + JS
  kernel instead of two ``scipy.stats.gaussian_kde`` and ``jensenshannon``.

Metrics outside SURVEY.md section 8 (TNR@TPR, AUROC, percentile scores, runtime/throughput,
classification) are "next" rows and are not built here: asking the factories for them raises
``ValueError`` naming the metric.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import ops


class UncertaintyEstimate:
    """Container for one model's uncertainty scores (reference ``evaluation.py:12-92``), kept as
    torch tensors on whatever device the model produced them."""

    def __init__(self, data: Union[np.ndarray, torch.Tensor, Tuple]):
        def numel(d):
            return d.numel() if isinstance(d, torch.Tensor) else d.size
        if isinstance(data, (np.ndarray, torch.Tensor)) and numel(data) == 0:
            raise ValueError("Cannot create UncertaintyEstimate from empty data")
        if isinstance(data, tuple) and any(numel(d) == 0 for d in data):
            raise ValueError("Cannot create UncertaintyEstimate from empty tuple data")
        self._t = self._to_tensor(data)
        if isinstance(self._t, tuple):
            shapes = [d.shape[0] for d in self._t]
            if len(set(shapes)) > 1:
                raise ValueError(
                    f"All arrays in tuple must have same first dimension, got shapes: {shapes}")

    @staticmethod
    def _to_tensor(data):
        if isinstance(data, torch.Tensor):
            return data.detach()
        if isinstance(data, np.ndarray):
            return torch.from_numpy(data)
        if isinstance(data, tuple):
            return tuple(UncertaintyEstimate._to_tensor(d) for d in data)
        raise TypeError(f"Unsupported data type: {type(data)}")

    @property
    def tensor(self):
        return self._t

    @property
    def data(self):
        """numpy view of the scores (device -> host copy on demand)."""
        if isinstance(self._t, tuple):
            return tuple(d.cpu().numpy() for d in self._t)
        return self._t.cpu().numpy()

    @property
    def dimensions(self) -> int:
        return len(self._t) if isinstance(self._t, tuple) else 1

    def flatten(self):
        if self.dimensions != 1:
            raise ValueError("Can only flatten 1D uncertainty estimates")
        return self._t.reshape(-1)

    def mean(self):
        if self.dimensions == 1:
            return float(self._t.double().mean())
        return float(torch.cat([d.reshape(-1).double() for d in self._t]).mean())


class EvaluationMetric(ABC):
    @abstractmethod
    def evaluate(self, model: nn.Module, id_data: tuple, ood_data: tuple) -> dict:
        pass

    @classmethod
    @abstractmethod
    def get_objectives(cls):
        pass

    @classmethod
    @abstractmethod
    def get_metrics(cls):
        pass

    @abstractmethod
    def get_name(cls):
        pass


class UncertaintyEvaluationMetric(EvaluationMetric):
    """Runs the two UQ forwards and hands the scores to the metric
    (reference ``evaluation.py:122-144``)."""

    def evaluate(self, model, id_data: tuple, ood_data: tuple) -> dict:
        model.eval()
        with torch.no_grad():
            _, id_scores = model(id_data[0], return_ue=True)
            _, ood_scores = model(ood_data[0], return_ue=True)
        result = self._evaluate_uncertainties(UncertaintyEstimate(id_scores),
                                              UncertaintyEstimate(ood_scores))
        return {k: float(v) for k, v in result.items()}

    def _evaluate_uncertainties(self, id_ue: UncertaintyEstimate, ood_ue: UncertaintyEstimate) -> dict:
        raise NotImplementedError


def _on_gpu(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("nnueehcs_b200 metrics run on CUDA only (no CPU fallback)")
        t = t.cuda()
    return t


class WassersteinEvaluation(UncertaintyEvaluationMetric):
    name = "wasserstein_distance"

    def _evaluate_uncertainties(self, id_ue, ood_ue) -> dict:
        if id_ue.dimensions != ood_ue.dimensions:
            raise ValueError("Uncertainty estimates must have the same dimensions")
        if id_ue.dimensions == 1:
            value = ops.wasserstein_1d(_on_gpu(id_ue.flatten()), _on_gpu(ood_ue.flatten()))
        else:
            value = float(np.mean([
                ops.wasserstein_1d(_on_gpu(a.reshape(-1)), _on_gpu(b.reshape(-1)))
                for a, b in zip(id_ue.tensor, ood_ue.tensor)]))
        return {self.name: value}

    @classmethod
    def get_objectives(cls):
        return [{"name": cls.name, "type": "maximize"}]

    @classmethod
    def get_metrics(cls):
        return [cls.name]

    def get_name(self):
        return self.name


class JensenShannonEvaluation(UncertaintyEvaluationMetric):
    name = "jensen_shannon_distance"

    def _evaluate_uncertainties(self, id_ue, ood_ue) -> dict:
        if id_ue.dimensions != ood_ue.dimensions:
            raise ValueError("Uncertainty estimates must have the same dimensions")
        return {self.name: self._average_js_distance(id_ue.tensor, ood_ue.tensor)}

    def _average_js_distance(self, p1, p2) -> float:
        if isinstance(p1, torch.Tensor) and (p1.dim() == 1 or (p1.dim() == 2 and p1.shape[1] == 1)):
            return self.pdf_jsd(p1.reshape(-1), p2.reshape(-1))
        raise ValueError("JensenShannonEvaluation: only 1-D (or [N, 1]) uncertainty scores are on "
                         "the accelerated path")

    def pdf_jsd(self, dist1, dist2, num_points=20000) -> float:
        if isinstance(dist1, np.ndarray):
            dist1 = torch.from_numpy(dist1)
        if isinstance(dist2, np.ndarray):
            dist2 = torch.from_numpy(dist2)
        return ops.kde_jsd(_on_gpu(dist1), _on_gpu(dist2), num_points)

    @classmethod
    def get_objectives(cls):
        return [{"name": cls.name, "type": "maximize"}]

    @classmethod
    def get_metrics(cls):
        return [cls.name]

    def get_name(self):
        return self.name


class MetricEvaluator:
    """Unified evaluator over several metrics (reference ``evaluation.py:666-697``)."""

    def __init__(self, metrics):
        self.metrics = metrics

    def evaluate(self, model: nn.Module, id_data: tuple, ood_data: tuple) -> dict:
        results = {}
        for metric in self.metrics:
            results.update(metric.evaluate(model, id_data, ood_data))
        return results

    def get_training_objectives(self):
        out = []
        for m in self.metrics:
            out.extend(m.get_instance_objectives() if hasattr(m, 'get_instance_objectives')
                       else m.get_objectives())
        return out

    def get_all_metrics(self):
        out = []
        for m in self.metrics:
            out.extend(m.get_instance_metrics() if hasattr(m, 'get_instance_metrics')
                       else m.get_metrics())
        return out


_DISTANCE_METRICS = {
    WassersteinEvaluation.name: WassersteinEvaluation,
    JensenShannonEvaluation.name: JensenShannonEvaluation,
}
# names the reference's factories also know (evaluation.py:700-812) but that are not on the hot path
_NEXT_ROWS = ("euclidean_distance", "percentile_classification", "tnr_at_tpr", "runtime",
              "uncertainty_estimating_runtime", "uncertainty_estimating_throughput",
              "base_model_throughput", "mean_score", "max_score", "percentile_score", "auroc",
              "max_memory_usage")


def _create_single_evaluator(metric_config: dict) -> EvaluationMetric:
    name = metric_config['name']
    if name in _DISTANCE_METRICS:
        return _DISTANCE_METRICS[name]()
    if name == 'wasserstein':  # spelling used by get_evaluator (evaluation.py:708)
        return WassersteinEvaluation()
    if name in _NEXT_ROWS:
        raise ValueError(f"metric '{name}' is not on the accelerated hot path (a 'next' row of "
                         "SURVEY.md section 8); use the reference's evaluator for it")
    raise ValueError(f"Invalid metric type: {name}")


def get_uncertainty_evaluator(metric_config) -> MetricEvaluator:
    """str | dict | list of those -> MetricEvaluator (reference ``evaluation.py:746-772``)."""
    configs = metric_config if isinstance(metric_config, list) else [metric_config]
    metrics = []
    for cfg in configs:
        if isinstance(cfg, str):
            cfg = {'name': cfg}
        metrics.append(_create_single_evaluator(cfg))
    return MetricEvaluator(metrics)


def get_evaluator(config) -> MetricEvaluator:
    """dict | list of dicts -> MetricEvaluator (reference ``evaluation.py:700-743``)."""
    configs = config if isinstance(config, list) else [config]
    return MetricEvaluator([_create_single_evaluator(c) for c in configs])
