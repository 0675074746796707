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

* the score consumers of SURVEY.md section 8f row 1 -- ``MeanScoreEvaluation``, ``MaxScoreEvaluation``,
  ``PercentileScoreEvaluation`` (``evaluation.py:292-381``), ``TNRatTPX`` (``:519-605``), ``AUROC``
  (``:607-635``) and ``PercentileBasedClassifier`` (``:637-662``) -- read one ``uq_score_metrics``
  result (two device sorts + binary searches) instead of numpy / sklearn / a Python loop over every
  unique score.

The host-side metrics around the same calls -- ``EuclideanEvaluation`` (``:205-227``),
``MaxMemoryUsageEvaluation`` and the runtime / throughput classes (``:383-516``; the last one is the
reference's own definition of UQ throughput) -- are mirrored as plain host code so that a
reference config moves over unchanged.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .classification import (PercentileBasedIdOodClassifier,
                             ReversedPercentileBasedIdOodClassifier)


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

    def _enqueue_uncertainties(self, id_ue, ood_ue):
        """1-D scores: the metric enqueued without a synchronisation (``ops.PendingMetric``) so
        that MetricEvaluator pays one synchronisation for all of its distance metrics; ``None``
        where the synchronous path has to run."""
        if id_ue.dimensions != ood_ue.dimensions or id_ue.dimensions != 1:
            return None
        return ops.wasserstein_1d_async(_on_gpu(id_ue.flatten()), _on_gpu(ood_ue.flatten()))

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
        if isinstance(p1, torch.Tensor) and isinstance(p2, torch.Tensor) and p1.dim() == 2:
            # [N, d > 1] scores: the reference's other branch (evaluation.py:263-266) -- the mean
            # over rows of scipy's jensenshannon(p1[i], p2[i]) on the raw score rows.  N x d scalar
            # work with no kernel behind it in the reference either: done on the host.
            a = p1.detach().cpu().numpy().astype(np.float64)
            b = p2.detach().cpu().numpy().astype(np.float64)
            if a.shape != b.shape:
                raise ValueError(f"operands could not be broadcast together with shapes "
                                 f"{a.shape} {b.shape}")
            return float(np.mean(_jensenshannon_rows(a, b)))
        raise ValueError("JensenShannonEvaluation: scores must be a 1-D or 2-D tensor")

    def _enqueue_uncertainties(self, id_ue, ood_ue):
        """See ``WassersteinEvaluation._enqueue_uncertainties``."""
        p1, p2 = id_ue.tensor, ood_ue.tensor
        if id_ue.dimensions != ood_ue.dimensions or not isinstance(p1, torch.Tensor) or \
                not (p1.dim() == 1 or (p1.dim() == 2 and p1.shape[1] == 1)) or \
                p1.numel() < 2 or p2.numel() < 2:
            return None
        return ops.kde_jsd_async(_on_gpu(p1.reshape(-1)), _on_gpu(p2.reshape(-1)), 20000)

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


def _jensenshannon_rows(p: np.ndarray, q: np.ndarray) -> np.ndarray:
    """``scipy.spatial.distance.jensenshannon(p[i], q[i])`` for every row (natural log): rows
    normalised to sum 1, m = (p + q) / 2, sqrt((KL(p || m) + KL(q || m)) / 2), with scipy's
    ``rel_entr`` conventions (0 where p == 0 and m >= 0, inf where a term is undefined)."""
    p = p / p.sum(axis=1, keepdims=True)
    q = q / q.sum(axis=1, keepdims=True)
    m = (p + q) / 2.0

    def rel_entr(x, y):
        out = np.full(x.shape, np.inf)
        pos = (x > 0) & (y > 0)
        out[pos] = x[pos] * np.log(x[pos] / y[pos])
        out[(x == 0) & (y >= 0)] = 0.0
        out[np.isnan(x) | np.isnan(y)] = np.nan
        return out

    js = rel_entr(p, m).sum(axis=1) + rel_entr(q, m).sum(axis=1)
    return np.sqrt(js / 2.0)


def _scores_1d(ue: UncertaintyEstimate) -> torch.Tensor:
    if ue.dimensions != 1:
        raise ValueError("score metrics need 1-D uncertainty estimates")
    return _on_gpu(ue.flatten())


class MeanScoreEvaluation(UncertaintyEvaluationMetric):
    """Mean ID score, a BO minimisation target (reference ``evaluation.py:292-316``)."""
    name = "mean_score"

    def _evaluate_uncertainties(self, id_ue, ood_ue) -> dict:
        if id_ue.dimensions != ood_ue.dimensions:
            raise ValueError("Uncertainty estimates must have the same dimensions")
        return {self.name: ops.score_metrics(_scores_1d(id_ue), _scores_1d(ood_ue))["mean_score"]}

    @classmethod
    def get_objectives(cls):
        return [{"name": cls.name, "type": "minimize"}]

    @classmethod
    def get_metrics(cls):
        return [cls.name]

    def get_name(self):
        return self.name


class MaxScoreEvaluation(UncertaintyEvaluationMetric):
    """Maximum ID score (reference ``evaluation.py:319-337``)."""
    name = "max_score"

    def _evaluate_uncertainties(self, id_ue, ood_ue) -> dict:
        return {self.name: ops.score_metrics(_scores_1d(id_ue), _scores_1d(ood_ue))["max_score"]}

    @classmethod
    def get_objectives(cls):
        return [{"name": cls.name, "type": "maximize"}]

    @classmethod
    def get_metrics(cls):
        return [cls.name]

    def get_name(self):
        return self.name


class PercentileScoreEvaluation(UncertaintyEvaluationMetric):
    """``np.percentile`` of the ID scores (reference ``evaluation.py:339-381``)."""
    name = "percentile_score"

    def __init__(self, percentile: float = 95.0):
        if not 0 <= percentile <= 100:
            raise ValueError(f"percentile must be between 0 and 100, got {percentile}")
        self.percentile = percentile

    @classmethod
    def from_config(cls, config: dict) -> 'PercentileScoreEvaluation':
        return cls(percentile=config.get('percentile', 95.0))

    def _evaluate_uncertainties(self, id_ue, ood_ue) -> dict:
        if id_ue.dimensions != ood_ue.dimensions:
            raise ValueError("Uncertainty estimates must have the same dimensions")
        r = ops.score_metrics(_scores_1d(id_ue), _scores_1d(ood_ue), percentile_q=self.percentile)
        return {self.name: r["percentile_score"]}

    @classmethod
    def get_objectives(cls):
        return [{"name": cls.name, "type": "minimize"}]

    @classmethod
    def get_metrics(cls):
        return [cls.name]

    def get_name(self):
        return self.name


class ClassificationMetric(EvaluationMetric):
    """Base of the classification-style metrics (reference ``evaluation.py:158-169``)."""

    def evaluate(self, model: nn.Module, id_data: tuple, ood_data: tuple) -> dict:
        with torch.no_grad():
            _, id_scores = model(id_data[0], return_ue=True)
            _, ood_scores = model(ood_data[0], return_ue=True)
        return self._evaluate_scores(id_scores, ood_scores)

    @abstractmethod
    def _evaluate_scores(self, id_scores: torch.Tensor, ood_scores: torch.Tensor) -> dict:
        pass


class TNRatTPX(ClassificationMetric):
    """True-negative rate at a target true-positive rate (reference ``evaluation.py:519-605``;
    its loop over every unique score is a closed form on the sorted arrays, see
    ``oracle/metrics_oracle.py:tnr_at_tpr``)."""

    def __init__(self, target_tpr: float, reversed: bool = False):
        if not 0 <= target_tpr <= 1:
            raise ValueError(f"target_tpr must be between 0 and 1, got {target_tpr}")
        self.target_tpr = target_tpr
        self.metric_name = 'tnr_at_tpr'
        self.reversed = reversed

    @classmethod
    def from_config(cls, config: dict) -> 'TNRatTPX':
        return cls(target_tpr=config['target_tpr'], reversed=config.get('reversed', False))

    def _evaluate_scores(self, id_scores, ood_scores) -> dict:
        r = ops.score_metrics(_on_gpu(id_scores), _on_gpu(ood_scores), target_tpr=self.target_tpr,
                              tnr_reversed=self.reversed)
        return {str(self): r["tnr_at_tpr"]}

    @classmethod
    def get_objectives(cls):
        return [{'name': 'tnr_at_tpr', 'type': 'maximize'}]

    @classmethod
    def get_metrics(cls):
        return ['tnr_at_tpr']

    def get_instance_objectives(self):
        return [{'name': self.metric_name, 'type': 'maximize'}]

    def get_instance_metrics(self):
        return [self.metric_name]

    def get_name(self):
        return f'{self.metric_name}{int(100*self.target_tpr)}'

    def __str__(self):
        return self.get_name()


class AUROC(ClassificationMetric):
    """Area under the ROC curve, OOD = positive class (reference ``evaluation.py:607-635``)."""
    name = "auroc"

    def _evaluate_scores(self, id_scores, ood_scores) -> dict:
        return {self.name: ops.score_metrics(_on_gpu(id_scores), _on_gpu(ood_scores))["auroc"]}

    @classmethod
    def get_objectives(cls):
        return [{'name': 'auroc', 'type': 'maximize'}]

    @classmethod
    def get_metrics(cls):
        return ['auroc']

    def get_name(self):
        return self.name


class PercentileBasedClassifier(ClassificationMetric):
    """Sensitivity / specificity of the percentile-threshold classifier (reference
    ``evaluation.py:637-662``; ``reversed`` negates the scores first)."""

    def __init__(self, percentile: float, reversed: bool = False):
        self._classifier = PercentileBasedIdOodClassifier(percentile)
        self.reversed = reversed

    def _evaluate_scores(self, id_scores, ood_scores) -> dict:
        r = ops.score_metrics(_on_gpu(id_scores), _on_gpu(ood_scores),
                              classifier_percentile=self._classifier.percentile,
                              classifier_reversed=self.reversed)
        return {k: r[k] for k in self.get_metrics()}

    @classmethod
    def get_objectives(cls):
        return [{'name': 'sensitivity', 'type': 'maximize'},
                {'name': 'specificity', 'type': 'maximize'}]

    @classmethod
    def get_metrics(cls):
        return ['sensitivity', 'specificity']

    def get_name(self):
        suffix = f'_{int(100*self._classifier.percentile)}'
        if self.reversed:
            suffix = f'_reversed{suffix}'
        return f'percentile_classification{suffix}'


# ------------------------------------------------------------------------------------------------
# Host-side metrics around the same calls (reference ``evaluation.py:205-227, 383-516``): nothing to
# accelerate, but the reference's configs name them (``uncertainty_estimating_throughput`` is the
# BO objective of examples/bo_driven/config*.yaml, and the reference's own definition of UQ
# throughput), so a config moves over unchanged.  Same names, result keys, objectives and timing
# protocol: wall clock around each call with a device synchronisation before the clock stops.
# ------------------------------------------------------------------------------------------------

class EuclideanEvaluation(UncertaintyEvaluationMetric):
    """Mean L2 distance between paired ID / OOD score rows (reference ``evaluation.py:205-227``)."""
    name = "euclidean_distance"

    def _evaluate_uncertainties(self, id_ue, ood_ue) -> dict:
        if id_ue.dimensions != ood_ue.dimensions:
            raise ValueError("Uncertainty estimates must have the same dimensions")
        diff = np.asarray(id_ue.data) - np.asarray(ood_ue.data)
        return {self.name: float(np.mean(np.sqrt(np.sum(diff ** 2, axis=-1))))}

    @classmethod
    def get_objectives(cls):
        return [{"name": cls.name, "type": "maximize"}]

    @classmethod
    def get_metrics(cls):
        return [cls.name]

    def get_name(self):
        return self.name


def _device_sync() -> None:
    if torch.cuda.is_available():
        torch.cuda.synchronize()


class MaxMemoryUsageEvaluation(EvaluationMetric):
    """Peak torch-allocator bytes (MiB) of one UQ forward over ID + OOD inputs (reference
    ``evaluation.py:383-411``).  The fused kernels take their workspaces from torch's caching
    allocator, so the number includes them."""
    name = "max_memory_usage"

    def evaluate(self, model, id_data: tuple, ood_data: tuple) -> dict:
        import gc
        model.eval()
        with torch.no_grad():
            torch.cuda.empty_cache()
            gc.collect()
            torch.cuda.reset_peak_memory_stats()
            model(torch.cat([id_data[0], ood_data[0]]), return_ue=True)
            peak = torch.cuda.max_memory_allocated()
        return {self.name: peak / float(1 << 20)}

    def get_objectives(self):
        return [{"name": self.name, "type": "minimize"}]

    def get_metrics(self):
        return [self.name]

    def get_name(self):
        return self.name


class RuntimeEvaluation(EvaluationMetric):
    """``num_warmup`` untimed + ``num_trials`` timed calls on the concatenated ID + OOD inputs
    (reference ``evaluation.py:413-461``); subclasses choose the call."""
    name = "runtime"

    def __init__(self, num_trials: int = 20, num_warmup: int = 5):
        self.num_trials = num_trials
        self.num_warmup = num_warmup

    @classmethod
    def from_config(cls, config: dict):
        return cls(num_trials=config.get('trials', 20), num_warmup=config.get('warmup', 5))

    def evaluate(self, model, id_data: tuple, ood_data: tuple) -> dict:
        raise NotImplementedError("Cannot call evaluate on base class")

    def _call(self, model, data):
        raise NotImplementedError

    def _seconds_per_call(self, model, id_data, ood_data) -> np.ndarray:
        import time
        model.eval()
        data = torch.cat([id_data[0], ood_data[0]])
        seconds = np.zeros(self.num_trials)
        with torch.no_grad():
            for _ in range(self.num_warmup):
                self._call(model, data)
            for i in range(self.num_trials):
                t0 = time.time()
                self._call(model, data)
                _device_sync()
                seconds[i] = time.time() - t0
        return seconds

    def _runtime(self, model, id_data, ood_data) -> dict:
        s = self._seconds_per_call(model, id_data, ood_data)
        return {'runtime': float(np.mean(s)), 'runtime_std': float(np.std(s))}

    @classmethod
    def get_objectives(cls):
        return [{"name": cls.name, "type": "minimize"}]

    @classmethod
    def get_metrics(cls):
        return [cls.name, 'runtime_std']

    def get_name(self):
        return self.name


class BaseModelRuntimeEvaluation(RuntimeEvaluation):
    name = "base_model_runtime"

    def _call(self, model, data):
        return model(data)

    def evaluate(self, model, id_data: tuple, ood_data: tuple) -> dict:
        return self._runtime(model, id_data, ood_data)


class UncertaintyEstimatingRuntimeEvaluation(RuntimeEvaluation):
    name = "uncertainty_estimating_runtime"

    def _call(self, model, data):
        return model(data, return_ue=True)

    def evaluate(self, model, id_data: tuple, ood_data: tuple) -> dict:
        return self._runtime(model, id_data, ood_data)


class BaseModelThroughputEvaluation(RuntimeEvaluation):
    """Samples per second of the plain forward: mean and std over the trials of
    ``total_samples / seconds`` (reference ``evaluation.py:475-492``)."""
    name = "base_model_throughput"

    def _call(self, model, data):
        return model(data)

    def evaluate(self, model, id_data: tuple, ood_data: tuple) -> dict:
        seconds = self._seconds_per_call(model, id_data, ood_data)
        rate = (id_data[0].shape[0] + ood_data[0].shape[0]) / seconds
        return {self.name: float(np.mean(rate)), 'throughput_std': float(np.std(rate))}


class UncertaintyEstimatingThroughputEvaluation(BaseModelThroughputEvaluation):
    """The reference's own definition of UQ throughput (``evaluation.py:494-516``): samples per
    second of ``model(x, return_ue=True)``."""
    name = "uncertainty_estimating_throughput"

    def _call(self, model, data):
        return model(data, return_ue=True)

    @classmethod
    def get_objectives(cls):
        return [{"name": cls.name, "type": "maximize"}]

    @classmethod
    def get_metrics(cls):
        return [cls.name]

    @classmethod
    def get_name(cls):
        return cls.name


class MetricEvaluator:
    """Unified evaluator over several metrics (reference ``evaluation.py:666-697``).

    The reference calls ``metric.evaluate(model, id, ood)`` per metric, and every one of them runs
    the two UQ forwards again (2 forwards per metric).  With
    ``share_forward=True`` (default) the forwards run once per model mode and every score-based
    metric reads the same device-resident scores; ``share_forward=False`` restores the
    reference's call pattern (for MC dropout that also means fresh masks per metric)."""

    def __init__(self, metrics, share_forward: bool = True):
        self.metrics = metrics
        self.share_forward = share_forward

    def evaluate(self, model: nn.Module, id_data: tuple, ood_data: tuple) -> dict:
        results = {}
        cache = {}   # model.training -> (id_scores, ood_scores)
        for metric in self.metrics:
            score_based = isinstance(metric, (UncertaintyEvaluationMetric, ClassificationMetric))
            if not (self.share_forward and score_based):
                results.update(metric.evaluate(model, id_data, ood_data))
                continue
            if isinstance(metric, UncertaintyEvaluationMetric):
                model.eval()   # as UncertaintyEvaluationMetric.evaluate does (evaluation.py:133)
            key = bool(getattr(model, "training", False))
            if key not in cache:
                with torch.no_grad():
                    _, id_scores = model(id_data[0], return_ue=True)
                    _, ood_scores = model(ood_data[0], return_ue=True)
                cache[key] = (id_scores, ood_scores)
            id_scores, ood_scores = cache[key]
            if isinstance(metric, UncertaintyEvaluationMetric):
                id_ue, ood_ue = UncertaintyEstimate(id_scores), UncertaintyEstimate(ood_scores)
                # the distance metrics are only ENQUEUED here (one launch each, no
                # synchronisation); they are read after the loop behind one synchronisation
                pending = (metric._enqueue_uncertainties(id_ue, ood_ue)
                           if hasattr(metric, "_enqueue_uncertainties") else None)
                if pending is not None:
                    results[metric.get_name()] = pending
                    continue
                r = metric._evaluate_uncertainties(id_ue, ood_ue)
                results.update({k: float(v) for k, v in r.items()})
            else:
                results.update(metric._evaluate_scores(id_scores, ood_scores))
        return {k: (v.result() if isinstance(v, ops.PendingMetric) else v)
                for k, v in results.items()}

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
    EuclideanEvaluation.name: EuclideanEvaluation,
}


def _create_single_evaluator(metric_config: dict) -> EvaluationMetric:
    """reference ``evaluation.py:774-812``"""
    name = metric_config['name']
    if name in _DISTANCE_METRICS:
        return _DISTANCE_METRICS[name]()
    if name == 'percentile_classification':
        cls = (ReversedPercentileBasedIdOodClassifier if metric_config.get('reversed', False)
               else PercentileBasedIdOodClassifier)
        return cls(metric_config['threshold'])
    if name == 'tnr_at_tpr':
        return TNRatTPX(metric_config['target_tpr'], metric_config.get('reversed', False))
    if name == 'mean_score':
        return MeanScoreEvaluation()
    if name == 'max_score':
        return MaxScoreEvaluation()
    if name == 'percentile_score':
        return PercentileScoreEvaluation.from_config(metric_config)
    if name == 'auroc':
        return AUROC()
    if name == 'runtime':
        kwargs = {}
        if 'trials' in metric_config:
            kwargs['num_trials'] = metric_config['trials']
        if 'warmup' in metric_config:
            kwargs['num_warmup'] = metric_config['warmup']
        return BaseModelRuntimeEvaluation(**kwargs)
    if name == 'uncertainty_estimating_runtime':
        return UncertaintyEstimatingRuntimeEvaluation()
    if name == 'uncertainty_estimating_throughput':
        return UncertaintyEstimatingThroughputEvaluation.from_config(metric_config)
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
    """dict | list of dicts -> MetricEvaluator (reference ``evaluation.py:700-743``: note its own
    spellings -- 'wasserstein', and 'percentile_classification' -> PercentileBasedClassifier)."""
    configs = config if isinstance(config, list) else [config]
    metrics = []
    for c in configs:
        name = c['name']
        if name == 'wasserstein':
            metrics.append(WassersteinEvaluation())
        elif name == 'percentile_classification':
            metrics.append(PercentileBasedClassifier(c['threshold'], c.get('reversed', False)))
        elif name == 'tnr_at_tpr':
            metrics.append(TNRatTPX.from_config(c))
        elif name == 'runtime':
            metrics.append(BaseModelRuntimeEvaluation.from_config(c))
        elif name == 'uncertainty_estimating_runtime':
            metrics.append(UncertaintyEstimatingRuntimeEvaluation.from_config(c))
        elif name == 'base_model_throughput':
            metrics.append(BaseModelThroughputEvaluation.from_config(c))
        elif name == 'uncertainty_estimating_throughput':
            metrics.append(UncertaintyEstimatingThroughputEvaluation.from_config(c))
        elif name == 'max_memory_usage':
            metrics.append(MaxMemoryUsageEvaluation())
        elif name in ('mean_score', 'max_score', 'percentile_score', 'auroc'):
            metrics.append(_create_single_evaluator(c))
        # other names: skipped, as the reference's get_evaluator does (evaluation.py:700-743)
    return MetricEvaluator(metrics)
