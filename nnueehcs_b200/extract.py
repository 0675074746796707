"""Walk a built ``nn.Sequential`` and describe it to the native library.

The reference never inspects its networks -- the wrappers just call ``nn.Sequential.__call__``
K times (nnueehcs/models.py:103,156-158).  The fused op needs the static structure instead:
a list of blocks ``Linear [-> BatchNorm1d(eval)] [-> ReLU] [-> Dropout]``, which is exactly what
``model_builder.build_network`` (nnueehcs/model_builder.py:30-73) emits for the MLP configs and
what ``MCDropoutModelBuilder._add_dropout`` (:254-263) turns them into.  Anything outside that
vocabulary (e.g. the Conv2d fixtures of the reference's builder tests) is a ``ValueError`` naming
the layer -- there is no silent CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn


@dataclass
class Block:
    linear: nn.Linear
    bn: Optional[nn.BatchNorm1d] = None
    relu: bool = False
    dropout: bool = False


def split_blocks(net: nn.Sequential) -> List[Block]:
    if not isinstance(net, nn.Sequential):
        raise ValueError(f"fused UQ forward needs an nn.Sequential, got {type(net).__name__}")
    blocks: List[Block] = []
    for idx, m in enumerate(net):
        if isinstance(m, nn.Linear):
            blocks.append(Block(linear=m))
            continue
        if not isinstance(m, (nn.BatchNorm1d, nn.ReLU, nn.Dropout)):
            raise ValueError(
                f"layer {idx}: {type(m).__name__} is not supported by the fused UQ forward "
                "(supported: Linear, BatchNorm1d, ReLU, Dropout)")
        if not blocks:
            raise ValueError(f"layer {idx} ({type(m).__name__}) precedes the first Linear")
        cur = blocks[-1]
        if isinstance(m, nn.BatchNorm1d):
            if cur.bn is not None or cur.relu or cur.dropout:
                raise ValueError(f"layer {idx}: BatchNorm1d must directly follow its Linear")
            if not m.track_running_stats or m.running_mean is None:
                raise ValueError(f"layer {idx}: BatchNorm1d without running statistics is not "
                                 "supported in inference")
            if m.num_features != cur.linear.out_features:
                raise ValueError(f"layer {idx}: BatchNorm1d width mismatch")
            cur.bn = m
        elif isinstance(m, nn.ReLU):
            if cur.relu or cur.dropout:
                raise ValueError(f"layer {idx}: unexpected ReLU position")
            cur.relu = True
        elif isinstance(m, nn.Dropout):
            if cur.dropout:
                raise ValueError(f"layer {idx}: two Dropouts in a row")
            cur.dropout = True
    if not blocks:
        raise ValueError("network has no Linear layer")
    for a, b in zip(blocks[:-1], blocks[1:]):
        if a.linear.out_features != b.linear.in_features:
            raise ValueError("consecutive Linear layers do not chain")
    return blocks


def structure_signature(blocks: Sequence[Block]):
    return tuple((b.linear.in_features, b.linear.out_features, b.bn is not None, b.relu, b.dropout)
                 for b in blocks)


def dropout_widths(blocks: Sequence[Block]) -> List[int]:
    """Width of the activation each Dropout acts on, in module order (mask layout)."""
    return [b.linear.out_features for b in blocks if b.dropout]


def _f32_cuda(t: Optional[torch.Tensor], device, keep: list) -> Optional[torch.Tensor]:
    """float32 contiguous view on `device`; the library reads through raw pointers, so the tensor
    is pinned in `keep` until the (synchronising) pack call returns."""
    if t is None:
        return None
    t = t.detach()
    if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    keep.append(t)
    return t


def tensors_version(nets: Sequence[nn.Sequential]):
    """Cache key of the packed weights: for every parameter and buffer its Python identity, storage
    address and in-place version -- plus the module structure (a ReLU swapped out, a Dropout
    removed or its ``p`` changed alters what is packed without touching a tensor).  A dtype or
    device change moves the storage, so the address covers those.  What no cheap key can see --
    ``param.data.mul_()`` (every ``.data`` is a fresh alias with its own version counter) or a
    write through a raw pointer -- is what ``WrappedModelBase.invalidate_packed()`` is for.

    This runs on every forward: one flat pass over the modules' own ``_parameters`` / ``_buffers``
    dicts (the networks are flat ``nn.Sequential``s of leaf modules, ``split_blocks``) -- the
    ``named_modules()`` / ``parameters()`` generators it replaced took 3.3 ms per call on a
    32-member 6 x 128 ensemble (1408 tensors), 12 % of that model's fused forward; this takes 1.1."""
    key = []
    add = key.append
    for net in nets:
        for mod in net._modules.values():
            add(id(mod))
            add(getattr(mod, "p", None))
            for t in mod._parameters.values():
                if t is not None:
                    add(id(t)), add(t._version), add(t.data_ptr())
            for t in mod._buffers.values():
                if t is not None:
                    add(id(t)), add(t._version), add(t.data_ptr())
    return tuple(key)
