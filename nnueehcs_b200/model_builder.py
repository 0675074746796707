"""Host-side mirror of ``nnueehcs/model_builder.py``: YAML architecture list -> ``nn.Sequential``
-> UQ wrapper.  Same class names, constructor arguments and ``build()`` / ``get_info()``
behaviour as the reference, so a model described by the reference's YAML configs is a drop-in;
the wrappers it returns are this package's (``models.py``), whose eval-mode forward is the fused
B200 op.

Behavioural notes checked against the reference:
* ``EnsembleModelBuilder`` seeds member ``i`` with ``torch.manual_seed(42 + i)``
  (model_builder.py:224-231) and never forwards ``vectorize``.
* ``MCDropoutModelBuilder`` inserts a Dropout immediately before every Linear/Conv2d of
  ``descr[1:-1]`` (model_builder.py:254-263).  The reference re-uses one dict object for all of
  them, so after the first ``build_network`` pop its later Dropouts are constructed with the
  default ``p``; ``MCDropoutModel.__init__`` then overwrites every ``p`` with
  ``dropout_percent`` anyway (models.py:132-134).  Here every Dropout is built with
  ``dropout_percent`` directly -- the same run-time behaviour.
* ``DeltaUQMLPModelBuilder`` doubles the first Linear's ``in_features`` once
  (model_builder.py:174-188) and requires ``estimator`` and ``anchored_batch_size`` keys.
"""
from __future__ import annotations

import copy
import types

import torch.nn

from .models import (DeltaUQMLP, EnsembleModel, KDEMLPModel, KNNKDEMLPModel, MCDropoutModel,
                     MLPModel, PAGERMLP)


class LayerBuilder:
    """Callable ``builder(name, *args, **kwargs) -> nn.Module`` that looks ``name`` up in an ordered
    list of namespaces (``torch.nn.__dict__`` by default; earlier namespaces shadow later ones).
    Same call / ``add_namespace`` interface as the reference's class of this name."""

    def __init__(self, *namespaces):
        self._spaces = list(namespaces)

    def _lookup(self, name):
        for space in self._spaces:
            if name in space:
                return space[name]
        raise KeyError(name)

    def __call__(self, name, *args, **kwargs):
        try:
            return self._lookup(name)(*args, **kwargs)
        except Exception as err:  # keep the exception type, add what was being built
            raise type(err)(str(err), name, args, kwargs) from err

    def add_namespace(self, namespace, index=-1):
        """``index >= 0``: insert at that search position; otherwise the new namespace is searched
        first."""
        self._spaces.insert(index if index >= 0 else 0, namespace)


def build_network(architecture, builder=LayerBuilder(torch.nn.__dict__)):
    """``[{LayerName: {args: [...], **kwargs}}, ...]`` -> ``nn.Sequential`` (the YAML format of
    examples/*/config.yaml; each list entry is a single-key dict).  The caller's description is
    left untouched."""
    modules = []
    for entry in architecture:
        if len(entry) != 1:
            raise AssertionError(f"one layer per list entry expected, got keys {list(entry)}")
        (layer_name, spec), = entry.items()
        options = copy.deepcopy(spec) if spec else {}
        positional = options.pop("args", [])
        modules.append(builder(layer_name, *positional, **options))
    return torch.nn.Sequential(*modules)


class InfoGrabbBase:
    """What the drivers ask about an architecture description: kind, depth, input width.  The
    width lives in the first layer's first positional argument (``Linear`` / ``Conv2d``)."""
    first_layer = None   # set by the two concrete grabbers

    def __init__(self, descr):
        self.descr = descr

    def num_layers(self):
        return len(self.descr)

    def is_mlp(self):
        return self.first_layer == 'Linear'

    def is_cnn(self):
        return self.first_layer == 'Conv2d'

    def num_inputs(self):
        return self.descr[0][self.first_layer]['args'][0]

    def set_num_inputs(self, num_inputs):
        self.descr[0][self.first_layer]['args'][0] = num_inputs


class CNNInfoGrabber(InfoGrabbBase):
    first_layer = 'Conv2d'


class MLPInfoGrabber(InfoGrabbBase):
    first_layer = 'Linear'


class ModelInfo:
    @classmethod
    def get_info_grabber(cls, model_descr):
        grabber = CNNInfoGrabber if 'Conv2d' in model_descr[0] else MLPInfoGrabber
        return grabber(model_descr)


def _attach(info, **getters):
    """Bind ``get_<name>()`` accessors returning fixed values onto an info grabber."""
    for name, value in getters.items():
        setattr(info, f"get_{name}", types.MethodType(lambda self, _v=value: _v, info))


class ModelBuilder:
    def __init__(self, model_descr, **kwargs):
        self.model_descr = copy.deepcopy(model_descr)
        self.train_config = kwargs.get('train_config')

    def build(self):
        return build_network(self.model_descr)

    def update_info(self, info):
        return info

    def get_info(self):
        info = ModelInfo.get_info_grabber(self.model_descr)
        self.update_info(info)
        return info


class MLPModelBuilder(ModelBuilder):
    def build(self):
        return MLPModel(super().build(), train_config=self.train_config)


class DeltaUQMLPModelBuilder(ModelBuilder):
    def __init__(self, base_descr, duq_descr, **kwargs):
        super().__init__(base_descr, **kwargs)
        self.duq_descr = duq_descr
        self._updated = False

    def build(self):
        self.update_info(self.get_info())
        return DeltaUQMLP(super().build(), train_config=self.train_config, **self.duq_descr)

    def update_info(self, info):
        # both keys are mandatory in the reference (KeyError otherwise, model_builder.py:176-177)
        _attach(info, estimator=self.duq_descr['estimator'],
                batch_size=self.duq_descr['anchored_batch_size'])
        if not self._updated:
            self._updated = True
            info.set_num_inputs(2 * info.num_inputs())


class PAGERModelBuilder(ModelBuilder):
    def __init__(self, base_descr, pager_descr, **kwargs):
        super().__init__(base_descr, **kwargs)
        self.pager_descr = pager_descr
        self._updated = False

    def build(self):
        self.update_info(self.get_info())
        return PAGERMLP(super().build(), train_config=self.train_config, **self.pager_descr)

    def update_info(self, info):
        _attach(info, estimator=self.pager_descr['estimator'])
        if not self._updated:
            self._updated = True
            info.set_num_inputs(2 * info.num_inputs())


class EnsembleModelBuilder(ModelBuilder):
    def __init__(self, base_descr, ensemble_descr, **kwargs):
        super().__init__(base_descr, **kwargs)
        self.ensemble_descr = ensemble_descr

    def build(self):
        info = self.get_info()
        members = []
        for i in range(info.get_num_models()):
            torch.manual_seed(42 + i)
            members.append(build_network(self.model_descr))
        return EnsembleModel(members, train_config=self.train_config)

    def update_info(self, info):
        _attach(info, num_models=self.ensemble_descr['num_models'])


class MCDropoutModelBuilder(ModelBuilder):
    def __init__(self, base_descr, dropout_descr, **kwargs):
        super().__init__(base_descr, **kwargs)
        self.dropout_descr = dropout_descr

    def build(self):
        self.model_descr = self._add_dropout(self.model_descr, self.dropout_descr)
        return MCDropoutModel(build_network(self.model_descr), train_config=self.train_config,
                              **self.dropout_descr)

    def _add_dropout(self, model_descr, dropout_descr):
        p = dropout_descr['dropout_percent']
        out = [model_descr[0]]
        for layer in model_descr[1:-1]:
            if layer.get('Linear') or layer.get('Conv2d'):
                out.append({'Dropout': {'args': [p]}})
            out.append(layer)
        out.append(model_descr[-1])
        return out

    def update_info(self, info):
        _attach(info, num_samples=self.dropout_descr['num_samples'],
                dropout_percent=self.dropout_descr['dropout_percent'])


class KDEModelBuilder(ModelBuilder):
    def __init__(self, base_descr, kde_descr, **kwargs):
        super().__init__(base_descr, **kwargs)
        self.kde_descr = kde_descr

    def build(self):
        return KDEMLPModel(super().build(), **self.kde_descr, train_config=self.train_config)


class KNNKDEModelBuilder(ModelBuilder):
    def __init__(self, base_descr, knn_kde_descr, **kwargs):
        super().__init__(base_descr, **kwargs)
        self.knn_kde_descr = knn_kde_descr

    def build(self):
        return KNNKDEMLPModel(super().build(), **self.knn_kde_descr,
                              train_config=self.train_config)
