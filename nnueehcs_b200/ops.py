"""Thin Python layer over the C ABI: torch supplies device memory and streams, nothing else.

Every function here ends in a call into ``libnnueehcs_b200.so``; there is no torch
implementation of any of these ops and no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .extract import (Block, _f32_cuda, dropout_widths, split_blocks, structure_signature)

_PREC = {"fp32": _lib.PREC_FP32, "float32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16,
         "bfloat16": _lib.PREC_BF16, "fp32_ffma": _lib.PREC_FP32_FFMA}
_MODE = {"ensemble": _lib.MODE_ENSEMBLE, "mc_dropout": _lib.MODE_MC_DROPOUT,
         "delta_uq": _lib.MODE_DELTA_UQ, "pager": _lib.MODE_PAGER}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """``torch.cuda.device(dev)`` without the set/restore round trip when ``dev`` is already the
    current device (the usual case: one process per GPU) -- a few microseconds per call, which is
    a visible share of a 100-microsecond metric call."""
    __slots__ = ("idx", "prev")

    def __init__(self, device: torch.device):
        self.idx = device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is not None and cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(
            f"nnueehcs_b200: {what} must be a CUDA tensor -- the fused UQ forward runs on the GPU "
            "only (no CPU fallback)")


def launch_count() -> int:
    return int(_lib.load().uq_launch_count())


def reset_launch_count() -> None:
    _lib.load().uq_launch_count_reset()


class PackedModel:
    """Device-resident packed weights of K structurally identical MLPs (``uq_model_t``)."""

    def __init__(self, nets: Sequence[nn.Sequential], device: torch.device,
                 anchor_first: bool = True):
        """``anchor_first`` only matters for ``mode='delta_uq'`` / ``'pager'``: True (default) is the
        public ``deltauq`` package's network input ``cat([anchor, x - anchor])``; False is
        ``cat([x - anchor, anchor])``, the order ``uq_forward`` evaluates natively.  The two differ
        by a swap of the first Linear's column halves (``UQ_MODEL_ANCHOR_FIRST``); the other modes
        do not look at the flag."""
        lib = _lib.load()
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("nnueehcs_b200: models can only be packed on a CUDA device")
        all_blocks = [split_blocks(net) for net in nets]
        sig = structure_signature(all_blocks[0])
        for i, b in enumerate(all_blocks[1:], 1):
            if structure_signature(b) != sig:
                raise ValueError(f"ensemble member {i} has a different architecture than member 0")
        self.signature = sig
        self.n_members = len(nets)
        self.n_layers = len(sig)
        self.d_in = sig[0][0]
        self.d_out = sig[-1][1]
        self.dropout_widths = dropout_widths(all_blocks[0])
        self.device = device
        self.anchor_first = bool(anchor_first)
        keep: list = []
        descs = (_lib.LayerDesc * (self.n_members * self.n_layers))()
        for k, blocks in enumerate(all_blocks):
            for l, blk in enumerate(blocks):
                d = descs[k * self.n_layers + l]
                self._fill(d, blk, device, keep)

        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.uq_model_create_ex(C.byref(handle), self.n_members, self.n_layers, descs,
                                              _lib.MODEL_ANCHOR_FIRST if self.anchor_first else 0,
                                              _stream_ptr(device)))
        self._handle = handle
        self._lib = lib
        self.supports_bf16 = bool(lib.uq_model_supports_bf16(handle))
        self.bf16_reason = "" if self.supports_bf16 else lib.uq_last_error().decode()
        # True: precision='fp32' runs as the tensor-core split kernel; False: CUDA-core FFMA
        self.fp32_on_tensor_cores = bool(lib.uq_model_supports_fp32_tc(handle))
        self.fp32_tc_reason = ("" if self.fp32_on_tensor_cores
                               else lib.uq_last_error().decode())
        del keep  # uq_model_create synchronised its stream: our staging copies may go

    @staticmethod
    def _fill(d: "_lib.LayerDesc", blk: Block, device, keep) -> None:
        lin = blk.linear
        d.in_features, d.out_features = lin.in_features, lin.out_features
        w = _f32_cuda(lin.weight, device, keep)
        b = _f32_cuda(lin.bias, device, keep)
        d.weight = w.data_ptr()
        d.bias = b.data_ptr() if b is not None else None
        if blk.bn is not None:
            bn = blk.bn
            g = _f32_cuda(bn.weight, device, keep)
            bb = _f32_cuda(bn.bias, device, keep)
            d.bn_weight = g.data_ptr() if g is not None else None
            d.bn_bias = bb.data_ptr() if bb is not None else None
            d.bn_mean = _f32_cuda(bn.running_mean, device, keep).data_ptr()
            d.bn_var = _f32_cuda(bn.running_var, device, keep).data_ptr()
            d.bn_eps = float(bn.eps)
        d.relu = 1 if blk.relu else 0
        d.dropout = 1 if blk.dropout else 0

    def close(self) -> None:
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.uq_model_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _args(self, mode, precision, total_members, member_begin, member_count, dropout_p,
              dropout_active, seed, offset, masks, anchors, output, row_base=0
              ) -> "_lib.ForwardArgs":
        a = _lib.ForwardArgs()
        a.mode = _MODE[mode]
        if precision not in _PREC:
            raise ValueError(f"unknown precision {precision!r} (use 'fp32', 'bf16' or "
                             "'fp32_ffma')")
        a.precision = _PREC[precision]
        a.output = _lib.OUT_MOMENTS if output == "moments" else _lib.OUT_MEAN_STD
        a.member_begin = int(member_begin)
        a.member_count = int(total_members - member_begin if member_count is None else member_count)
        a.total_members = int(total_members)
        a.dropout_active = 1 if dropout_active else 0
        a.row_base = int(row_base)
        a.dropout_p = float(dropout_p)
        a.philox_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        a.philox_offset = int(offset) & 0xFFFFFFFFFFFFFFFF
        a.masks = masks.data_ptr() if masks is not None else None
        a.anchors = anchors.data_ptr() if anchors is not None else None
        return a

    def forward(self, x: torch.Tensor, mode: str, *, total_members: int, precision: str = "fp32",
                member_begin: int = 0, member_count: Optional[int] = None,
                dropout_p: float = 0.0, dropout_active: bool = True, seed: int = 0,
                offset: int = 0, masks: Optional[torch.Tensor] = None,
                anchors: Optional[torch.Tensor] = None,
                output: str = "mean_std", targets: Optional[torch.Tensor] = None,
                score_floor: Optional[torch.Tensor] = None, row_base: int = 0
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(mean, std) -- or (mean, M2) with ``output='moments'`` -- of shape ``[n, d_out]``.
        ``mode='pager'``: (mean over anchors of the swapped-role predictions, conformal score
        ``max_k |P[n, k] - targets[k]|`` raised to ``score_floor`` when given)."""
        _require_cuda(x, "x")
        d_x = self.d_in // 2 if mode in ("delta_uq", "pager") else self.d_in
        if x.dim() != 2 or x.shape[1] != d_x:
            raise ValueError(f"x must be [n, {d_x}], got {tuple(x.shape)}")
        if x.shape[0] == 0:
            raise ValueError("x has no rows")
        xf = x.detach()
        if xf.dtype != torch.float32 or not xf.is_contiguous():
            xf = xf.to(torch.float32).contiguous()
        if masks is not None:
            _require_cuda(masks, "masks")
            if masks.dtype != torch.uint8 or not masks.is_contiguous():
                raise ValueError("masks must be a contiguous uint8 tensor")
            need = sum(total_members * x.shape[0] * w for w in self.dropout_widths)
            if masks.numel() != need:
                raise ValueError(f"masks has {masks.numel()} bytes, expected {need}")
        if anchors is not None:
            _require_cuda(anchors, "anchors")
            anchors = anchors.detach().to(torch.float32).contiguous()
            if anchors.dim() != 2 or anchors.shape[1] != d_x or anchors.shape[0] < total_members:
                raise ValueError(f"anchors must be [>= {total_members}, {d_x}]")
        if mode == "pager":
            if targets is None:
                raise ValueError("PAGER needs the anchors' targets (anchors_Y)")
            _require_cuda(targets, "targets")
            targets = targets.detach().to(torch.float32).reshape(-1, self.d_out).contiguous()
            if targets.shape[0] < total_members:
                raise ValueError(f"targets must be [>= {total_members}, {self.d_out}]")
            if score_floor is not None:
                _require_cuda(score_floor, "score_floor")
                score_floor = score_floor.detach().to(torch.float32).contiguous()
                if tuple(score_floor.shape) != (x.shape[0], self.d_out):
                    raise ValueError(f"score_floor must be [{x.shape[0]}, {self.d_out}]")
        elif targets is not None or score_floor is not None:
            raise ValueError("targets / score_floor only apply to mode='pager'")
        if precision not in _PREC:
            raise ValueError(f"unknown precision {precision!r} (use 'fp32', 'bf16' or "
                             "'fp32_ffma')")
        count = int(total_members - member_begin if member_count is None else member_count)
        # the registered custom op (torch.ops.nnueehcs_b200.uq_forward, CUDA dispatch key only)
        return torch.ops.nnueehcs_b200.uq_forward(
            int(self._handle.value), xf, _MODE[mode], _PREC[precision],
            _lib.OUT_MOMENTS if output == "moments" else _lib.OUT_MEAN_STD, int(member_begin),
            count, int(total_members), bool(dropout_active), float(dropout_p),
            _as_i64(seed), _as_i64(offset), masks, anchors, int(self.d_out), targets, score_floor,
            int(row_base))

    def forward_into(self, x: torch.Tensor, mode: str, out0: torch.Tensor, out1: torch.Tensor, *,
                     total_members: int, precision: str = "fp32", member_begin: int = 0,
                     member_count: Optional[int] = None, dropout_p: float = 0.0,
                     dropout_active: bool = True, seed: int = 0, offset: int = 0,
                     masks: Optional[torch.Tensor] = None, anchors: Optional[torch.Tensor] = None,
                     output: str = "mean_std") -> None:
        """``forward`` writing into caller-owned ``[n, d_out]`` float32 tensors (e.g. the two halves
        of one exchange slab, ``distributed.KShard``) instead of allocating its outputs."""
        _require_cuda(x, "x")
        d_x = self.d_in // 2 if mode in ("delta_uq", "pager") else self.d_in
        if x.dim() != 2 or x.shape[1] != d_x or x.shape[0] == 0:
            raise ValueError(f"x must be [n >= 1, {d_x}], got {tuple(x.shape)}")
        for t in (out0, out1):
            _require_cuda(t, "out")
            if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != x.shape[0] * self.d_out:
                raise ValueError(f"outputs must be contiguous float32 [{x.shape[0]}, {self.d_out}]")
        xf = x.detach()
        if xf.dtype != torch.float32 or not xf.is_contiguous():
            xf = xf.to(torch.float32).contiguous()
        if anchors is not None:
            anchors = anchors.detach().to(torch.float32).contiguous()
        a = self._args(mode, precision, total_members, member_begin, member_count, dropout_p,
                       dropout_active, seed, offset, masks, anchors, output)
        _run_forward(self._lib, self._handle, xf, a, out0, out1)

    def forward_host(self, x_host: torch.Tensor, out0_host: torch.Tensor, out1_host: torch.Tensor,
                     mode: str, *, total_members: int, precision: str = "fp32",
                     dropout_p: float = 0.0, dropout_active: bool = True, seed: int = 0,
                     offset: int = 0, anchors: Optional[torch.Tensor] = None) -> None:
        """End-to-end call on HOST buffers (``uq_forward_host``): H2D, forward, D2H, sync."""
        for t in (x_host, out0_host, out1_host):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("forward_host takes contiguous float32 CPU tensors")
        a = self._args(mode, precision, total_members, 0, None, dropout_p, dropout_active, seed,
                       offset, None, anchors, "mean_std")
        with torch.cuda.device(self.device):
            _lib.check(self._lib.uq_forward_host(self._handle, x_host.data_ptr(), x_host.shape[0],
                                                 C.byref(a), out0_host.data_ptr(),
                                                 out1_host.data_ptr(), _stream_ptr(self.device)))


def philox_keep_masks(n: int, widths: Sequence[int], total_members: int, dropout_p: float, seed: int,
                      offset: int, device) -> torch.Tensor:
    """The keep-masks the native Philox path draws, in the injected-mask layout (uint8)."""
    lib = _lib.load()
    device = torch.device(device)
    total = sum(total_members * n * w for w in widths)
    out = torch.empty(total, dtype=torch.uint8, device=device)
    off = 0
    with torch.cuda.device(device):
        for l, w in enumerate(widths):
            _lib.check(lib.uq_philox_keep_masks(out.data_ptr() + off, n, w, total_members, l,
                                                float(dropout_p), int(seed), int(offset),
                                                _stream_ptr(device)))
            off += total_members * n * w
    return out


def moments_merge(means: torch.Tensor, m2s: torch.Tensor, counts: Sequence[float]
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Chan-merge ``[S, ...]`` shard moments into (mean, unbiased std) of shape ``[...]``."""
    _require_cuda(means, "means")
    _require_cuda(m2s, "m2s")
    means = means.contiguous()
    m2s = m2s.contiguous()
    if means.dtype != torch.float32 or m2s.dtype != torch.float32 or means.shape != m2s.shape:
        raise ValueError("means / m2s must be float32 tensors of the same shape")
    s = means.shape[0]
    if len(counts) != s:
        raise ValueError("one count per shard is required")
    return torch.ops.nnueehcs_b200.moments_merge(means, m2s, [float(c) for c in counts])


def moments_merge_strided(shards: torch.Tensor, counts: Sequence[float], output: str = "mean_std"
                          ) -> torch.Tensor:
    """Chan merge of ``shards`` = float32 ``[S, 2, L]`` (shard s: its means, then its M2 -- the
    receive buffer of the K-shard all-to-all) into a ``[2, L]`` tensor: (mean, unbiased std), or
    (mean, M2) with ``output='moments'``.  Shards whose count is 0 are skipped."""
    lib = _lib.load()
    _require_cuda(shards, "shards")
    if shards.dtype != torch.float32 or shards.dim() != 3 or shards.shape[1] != 2 \
            or not shards.is_contiguous():
        raise ValueError("shards must be a contiguous float32 [S, 2, L] tensor")
    s, _, length = shards.shape
    if len(counts) != s:
        raise ValueError("one count per shard is required")
    out = torch.empty((2, length), dtype=torch.float32, device=shards.device)
    cnt = (C.c_double * s)(*[float(c) for c in counts])
    base = shards.data_ptr()
    with torch.cuda.device(shards.device):
        _lib.check(lib.uq_moments_merge_ex(base, base + 4 * length, 2 * length, cnt, s, length,
                                           out[0].data_ptr(), out[1].data_ptr(),
                                           _lib.OUT_MOMENTS if output == "moments"
                                           else _lib.OUT_MEAN_STD, _stream_ptr(shards.device)))
    return out


def _flat_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    _require_cuda(t, what)
    t = t.detach().reshape(-1)
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


_WMETHOD = {"auto": _lib.WASSERSTEIN_AUTO, "sort": _lib.WASSERSTEIN_SORT,
            "binned": _lib.WASSERSTEIN_BINNED}


def wasserstein_1d(u: torch.Tensor, v: torch.Tensor, method: str = "auto") -> float:
    """``scipy.stats.wasserstein_distance(u, v)`` for float32 device samples."""
    u, v = _flat_f32(u, "u"), _flat_f32(v, "v")
    if u.numel() == 0 or v.numel() == 0:
        raise ValueError("Distribution can't be empty.")
    if method not in _WMETHOD:
        raise ValueError(f"unknown Wasserstein method {method!r} (auto, sort, binned)")
    return float(torch.ops.nnueehcs_b200.wasserstein_1d(u, v, _WMETHOD[method]))


def wasserstein_1d_info(u: torch.Tensor, v: torch.Tensor, method: str = "auto") -> dict:
    """Same value plus which method ran and how many values had to be sorted."""
    lib = _lib.load()
    u, v = _flat_f32(u, "u"), _flat_f32(v, "v")
    if u.numel() == 0 or v.numel() == 0:
        raise ValueError("Distribution can't be empty.")
    out, info = C.c_double(), (C.c_int64 * 3)()
    with torch.cuda.device(u.device):
        wsb = int(lib.uq_wasserstein_workspace_bytes(u.numel(), v.numel()))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=u.device)
        _lib.check(lib.uq_wasserstein_1d_ex(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(),
                                            _WMETHOD[method], C.byref(out), info, ws.data_ptr(),
                                            wsb, _stream_ptr(u.device)))
    return {"value": float(out.value), "method": "binned" if info[0] == 2 else "sort",
            "sorted_u": int(info[1]), "sorted_v": int(info[2])}


class PendingMetric:
    """A distribution metric enqueued on the current stream by ``wasserstein_1d_async`` /
    ``kde_jsd_async`` (``uq_*_enqueue``): nothing has been synchronised yet.  ``result()``
    synchronises that stream once (a no-op for every further pending metric of the same stream)
    and returns the float; the rare inputs the single-launch kernels do not cover (ambiguous
    bins, inf / NaN, a range beyond ~4000 bandwidths) run the synchronous call there."""
    __slots__ = ("kind", "u", "v", "num_points", "record", "ws", "wsb", "stream", "_value")

    def __init__(self, kind, u, v, num_points, record, ws, wsb, stream):
        self.kind, self.u, self.v, self.num_points = kind, u, v, num_points
        self.record, self.ws, self.wsb, self.stream = record, ws, wsb, stream
        self._value = None

    def result(self) -> float:
        if self._value is None:
            lib = _lib.load()
            self.stream.synchronize()
            out = C.c_double()
            u, v = self.u, self.v
            with _on_device(u.device):
                if self.kind == "wasserstein":
                    _lib.check(lib.uq_wasserstein_1d_finish(
                        u.data_ptr(), u.numel(), v.data_ptr(), v.numel(), self.record.data_ptr(),
                        C.byref(out), None, self.ws.data_ptr(), self.wsb, self.stream.cuda_stream))
                else:
                    _lib.check(lib.uq_kde_jsd_finish(
                        u.data_ptr(), u.numel(), v.data_ptr(), v.numel(), self.num_points,
                        self.record.data_ptr(), C.byref(out), None, self.ws.data_ptr(), self.wsb,
                        self.stream.cuda_stream))
            self._value = float(out.value)
            _free_records.append(self.record)
            self.u = self.v = self.ws = self.record = None     # release the buffers
        return self._value


_free_records: list = []     # pinned 256-byte result records, recycled by PendingMetric.result()
_async_ws: dict = {}         # (device index, stream) -> workspace shared by the enqueued metrics


def _metric_record() -> torch.Tensor:
    # pinned host memory is device-mapped under unified addressing
    if _free_records:
        return _free_records.pop()
    return torch.zeros(_lib.METRIC_RECORD_BYTES, dtype=torch.uint8, pin_memory=True)


def _shared_metric_ws(device: torch.device, stream, nbytes: int) -> torch.Tensor:
    """ONE workspace per (device, stream) for all enqueued metrics: calls on a stream run in order,
    each begins by zeroing its tables, and the only later reader (``uq_*_finish`` falling back to
    the synchronous path) recomputes everything after the stream has been synchronised -- so
    metrics in flight do not need a workspace each (0.8 GB at 50 M + 50 M values)."""
    key = (device.index, stream.cuda_stream)
    ws = _async_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _async_ws.pop(key, None)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _async_ws[key] = ws
    return ws


def wasserstein_1d_async(u: torch.Tensor, v: torch.Tensor) -> PendingMetric:
    """``wasserstein_1d(u, v)`` enqueued on the current stream, not synchronised
    (``uq_wasserstein_1d_enqueue``); ``.result()`` gives the float."""
    lib = _lib.load()
    u, v = _flat_f32(u, "u"), _flat_f32(v, "v")
    if u.numel() == 0 or v.numel() == 0:
        raise ValueError("Distribution can't be empty.")
    record = _metric_record()
    with _on_device(u.device):
        stream = torch.cuda.current_stream(u.device)
        wsb = int(lib.uq_wasserstein_workspace_bytes(u.numel(), v.numel()))
        ws = _shared_metric_ws(u.device, stream, wsb)
        _lib.check(lib.uq_wasserstein_1d_enqueue(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(),
                                                 record.data_ptr(), ws.data_ptr(), wsb,
                                                 stream.cuda_stream))
    return PendingMetric("wasserstein", u, v, 0, record, ws, wsb, stream)


def kde_jsd_async(u: torch.Tensor, v: torch.Tensor, num_points: int = 20000) -> PendingMetric:
    """``kde_jsd(u, v, num_points)`` enqueued on the current stream, not synchronised
    (``uq_kde_jsd_enqueue``); ``.result()`` gives the float."""
    lib = _lib.load()
    u, v = _flat_f32(u, "u"), _flat_f32(v, "v")
    record = _metric_record()
    with _on_device(u.device):
        stream = torch.cuda.current_stream(u.device)
        wsb = int(lib.uq_kde_jsd_workspace_bytes(u.numel(), v.numel(), int(num_points)))
        ws = _shared_metric_ws(u.device, stream, wsb)
        _lib.check(lib.uq_kde_jsd_enqueue(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(),
                                          int(num_points), record.data_ptr(), ws.data_ptr(), wsb,
                                          stream.cuda_stream))
    return PendingMetric("kde_jsd", u, v, int(num_points), record, ws, wsb, stream)


def sort_f32(x: torch.Tensor) -> torch.Tensor:
    """``x`` sorted ascending (a new float32 tensor): the radix sort behind the SORT Wasserstein
    method, the score metrics and the WINDOW KDE, on its own.  Ordering = the float32 bit patterns'
    (what ``np.sort`` gives, with ``-0.0`` before ``+0.0`` and NaNs last)."""
    lib = _lib.load()
    x = _flat_f32(x, "x")
    out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        wsb = int(lib.uq_sort_workspace_bytes(x.numel()))
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        _lib.check(lib.uq_sort_f32(x.data_ptr(), x.numel(), out.data_ptr(), ws.data_ptr(), wsb,
                                   _stream_ptr(x.device)))
    return out


_KMETHOD = {"auto": _lib.KDE_AUTO, "window": _lib.KDE_WINDOW, "moments": _lib.KDE_MOMENTS}


def kde_jsd(u: torch.Tensor, v: torch.Tensor, num_points: int = 20000, method: str = "auto"
            ) -> float:
    """``JensenShannonEvaluation.pdf_jsd(u, v, num_points)`` for float32 device samples."""
    u, v = _flat_f32(u, "u"), _flat_f32(v, "v")
    if method not in _KMETHOD:
        raise ValueError(f"unknown KDE method {method!r} (auto, window, moments)")
    return float(torch.ops.nnueehcs_b200.kde_jsd(u, v, int(num_points), _KMETHOD[method]))


def kde_jsd_info(u: torch.Tensor, v: torch.Tensor, num_points: int = 20000, method: str = "auto"
                 ) -> dict:
    """Same value plus which method ran."""
    lib = _lib.load()
    u, v = _flat_f32(u, "u"), _flat_f32(v, "v")
    out, used = C.c_double(), C.c_int32()
    with torch.cuda.device(u.device):
        wsb = int(lib.uq_kde_jsd_workspace_bytes(u.numel(), v.numel(), num_points))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=u.device)
        _lib.check(lib.uq_kde_jsd_ex(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(), num_points,
                                     _KMETHOD[method], C.byref(out), C.byref(used), ws.data_ptr(),
                                     wsb, _stream_ptr(u.device)))
    return {"value": float(out.value), "method": "moments" if used.value == 2 else "window"}


def kde_scott_bandwidth(m: int, d: int) -> float:
    """sklearn ``KernelDensity(bandwidth='scott').fit(X).bandwidth_`` for X of shape [m, d]."""
    return float(_lib.load().uq_kde_scott_bandwidth(int(m), int(d)))


def kde_silverman_bandwidth(m: int, d: int) -> float:
    """sklearn ``KernelDensity(bandwidth='silverman').fit(X).bandwidth_`` for X of shape [m, d]."""
    return float(_lib.load().uq_kde_silverman_bandwidth(int(m), int(d)))


def kde_density(fit: torch.Tensor, x: torch.Tensor, bandwidth: float) -> torch.Tensor:
    """``-exp(KernelDensity(bandwidth).fit(fit).score_samples(x))`` as a float64 [n] device tensor
    (KDEMLPModel.forward's uncertainty score, reference models.py:209-222)."""
    lib = _lib.load()
    _require_cuda(fit, "fit")
    _require_cuda(x, "x")
    if fit.dim() != 2 or x.dim() != 2 or fit.shape[1] != x.shape[1]:
        raise ValueError(f"fit [m, d] and x [n, d] must agree in d, got {tuple(fit.shape)} and "
                         f"{tuple(x.shape)}")
    if fit.shape[0] == 0 or x.shape[0] == 0:
        raise ValueError("kde_density needs at least one fitted row and one query row")
    fit = fit.detach().to(torch.float32).contiguous()
    x = x.detach().to(torch.float32).contiguous()
    n, m, d = x.shape[0], fit.shape[0], x.shape[1]
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        wsb = int(lib.uq_kde_density_workspace_bytes(n, m))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=x.device)
        _lib.check(lib.uq_kde_density(fit.data_ptr(), m, x.data_ptr(), n, d, float(bandwidth),
                                      out.data_ptr(), ws.data_ptr(), wsb, _stream_ptr(x.device)))
    return out


def score_metrics(id_scores: torch.Tensor, ood_scores: torch.Tensor, *, percentile_q: float = 95.0,
                  target_tpr: float = 0.95, tnr_reversed: bool = False,
                  classifier_percentile: float = 0.95, classifier_reversed: bool = False) -> dict:
    """Mean / max / percentile score, AUROC, TNR@TPR and the percentile classifier's rates from one
    pair of device sorts (``uq_score_metrics``)."""
    lib = _lib.load()
    a, b = _flat_f32(id_scores, "id_scores"), _flat_f32(ood_scores, "ood_scores")
    if a.numel() == 0 or b.numel() == 0:
        raise ValueError("score vectors must not be empty")
    req = _lib.ScoreRequest(float(percentile_q), float(target_tpr), float(classifier_percentile),
                            1 if tnr_reversed else 0, 1 if classifier_reversed else 0)
    out = _lib.ScoreResult()
    with torch.cuda.device(a.device):
        wsb = int(lib.uq_score_metrics_workspace_bytes(a.numel(), b.numel()))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=a.device)
        _lib.check(lib.uq_score_metrics(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(),
                                        C.byref(req), C.byref(out), ws.data_ptr(), wsb,
                                        _stream_ptr(a.device)))
    return {k: float(getattr(out, k)) for k, _ in _lib.ScoreResult._fields_}


# ------------------------------------------------------------------------------------------------
# per-rank steps of the sharded metrics (strung together by nnueehcs_b200.distributed)
# ------------------------------------------------------------------------------------------------

def key_bins() -> int:
    return int(_lib.load().uq_key_bins())


def sample_stats(x: torch.Tensor) -> Tuple[float, float, float, float]:
    """(min, max, mean, M2) of a float32 device shard, float64 arithmetic."""
    lib = _lib.load()
    x = _flat_f32(x, "x")
    out = (C.c_double * 4)()
    with torch.cuda.device(x.device):
        wsb = int(lib.uq_sample_stats_workspace_bytes())
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        _lib.check(lib.uq_sample_stats(x.data_ptr(), x.numel(), out, ws.data_ptr(), wsb,
                                       _stream_ptr(x.device)))
    return float(out[0]), float(out[1]), float(out[2]), float(out[3])


def kde_grid_accumulate(x: torch.Tensor, lo: float, hi: float, bandwidth: float,
                        grid: torch.Tensor) -> None:
    """grid[j] += sum_i exp(-((x_i - g_j) / bandwidth)^2 / 2), g = linspace(lo, hi, len(grid))."""
    lib = _lib.load()
    x = _flat_f32(x, "x")
    _require_cuda(grid, "grid")
    if grid.dtype != torch.float64 or not grid.is_contiguous() or grid.dim() != 1:
        raise ValueError("grid must be a contiguous 1-D float64 tensor")
    with torch.cuda.device(x.device):
        wsb = int(lib.uq_kde_grid_workspace_bytes(x.numel()))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=x.device)
        _lib.check(lib.uq_kde_grid_accumulate(x.data_ptr(), x.numel(), float(lo), float(hi),
                                              float(bandwidth), grid.numel(), grid.data_ptr(),
                                              ws.data_ptr(), wsb, _stream_ptr(x.device)))


def jsd_from_grids(grids: torch.Tensor) -> float:
    """Jensen-Shannon distance of two raw kernel-sum vectors, ``grids`` = float64 [2, G] on device."""
    lib = _lib.load()
    _require_cuda(grids, "grids")
    if grids.dtype != torch.float64 or grids.dim() != 2 or grids.shape[0] != 2:
        raise ValueError("grids must be float64 [2, G]")
    grids = grids.contiguous()
    out = C.c_double()
    with torch.cuda.device(grids.device):
        _lib.check(lib.uq_jsd_from_grids(grids.data_ptr(), grids.shape[1], C.byref(out),
                                         _stream_ptr(grids.device)))
    return float(out.value)


def key_histogram(x: torch.Tensor) -> torch.Tensor:
    """int64 [key_bins()] counts of ``x`` per coarse order-preserving key bin (on x's device)."""
    lib = _lib.load()
    x = _flat_f32(x, "x")
    hist = torch.zeros(key_bins(), dtype=torch.int32, device=x.device)
    if x.numel():
        with torch.cuda.device(x.device):
            _lib.check(lib.uq_key_histogram(x.data_ptr(), x.numel(), hist.data_ptr(),
                                            _stream_ptr(x.device)))
    return hist.to(torch.int64)


def partition_by_bin(x: torch.Tensor, bin_to_part: torch.Tensor, part_counts: Sequence[int]
                     ) -> torch.Tensor:
    """``x`` regrouped into ``len(part_counts)`` consecutive segments (segment p holds the values
    whose key bin maps to part p; order inside a segment is arbitrary)."""
    lib = _lib.load()
    x = _flat_f32(x, "x")
    _require_cuda(bin_to_part, "bin_to_part")
    if bin_to_part.dtype != torch.uint8 or bin_to_part.numel() != key_bins():
        raise ValueError("bin_to_part must be uint8 [key_bins()]")
    out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    starts, acc = [], 0
    for c in part_counts:
        starts.append(acc)
        acc += int(c)
    if acc != x.numel():
        raise ValueError("part_counts do not add up to the number of values")
    cursors = torch.tensor(starts, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.uq_partition_by_bin(x.data_ptr(), x.numel(),
                                           bin_to_part.contiguous().data_ptr(), len(part_counts),
                                           out.data_ptr(), cursors.data_ptr(),
                                           _stream_ptr(x.device)))
    return out


def bin_moments(x: torch.Tensor, tables: torch.Tensor, row: int) -> None:
    """Adds shard ``x`` to ``tables[row]`` (counts per key bin) and ``tables[row + 1]`` (sum of the
    low 18 key bits per bin); ``tables`` = int64 [4, key_bins()] on x's device."""
    lib = _lib.load()
    x = _flat_f32(x, "x")
    _require_cuda(tables, "tables")
    if tables.dtype != torch.int64 or tuple(tables.shape) != (4, key_bins()) \
            or not tables.is_contiguous():
        raise ValueError("tables must be a contiguous int64 [4, key_bins()] tensor")
    if x.numel() == 0:
        return
    with torch.cuda.device(x.device):
        _lib.check(lib.uq_bin_moments(x.data_ptr(), x.numel(), tables[row].data_ptr(),
                                      tables[row + 1].data_ptr(), _stream_ptr(x.device)))


def wasserstein_from_bins(tables: torch.Tensor, nu_total: int, nv_total: int) -> dict:
    """Resolved part of the Wasserstein integral from complete (all-reduced) bin tables."""
    lib = _lib.load()
    _require_cuda(tables, "tables")
    dev = tables.device
    flags = torch.empty(key_bins(), dtype=torch.uint8, device=dev)
    out = (C.c_double * 4)()
    with torch.cuda.device(dev):
        wsb = int(lib.uq_wasserstein_workspace_bytes(1, 1))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(lib.uq_wasserstein_from_bins(tables.data_ptr(), int(nu_total), int(nv_total),
                                                flags.data_ptr(), out, ws.data_ptr(), wsb,
                                                _stream_ptr(dev)))
    return {"resolved": float(out[0]), "amb_u": int(out[1]), "amb_v": int(out[2]),
            "nonfinite": int(out[3]), "flags": flags}


def compact_flagged(x: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """The values of ``x`` whose key bin is flagged (order arbitrary)."""
    lib = _lib.load()
    x = _flat_f32(x, "x")
    if x.numel() == 0:
        return x
    out = torch.empty_like(x)
    cnt = C.c_int64()
    with torch.cuda.device(x.device):
        ws = torch.empty(256, dtype=torch.uint8, device=x.device)
        _lib.check(lib.uq_compact_flagged(x.data_ptr(), x.numel(), flags.data_ptr(),
                                          out.data_ptr(), C.byref(cnt), ws.data_ptr(), 256,
                                          _stream_ptr(x.device)))
    return out[:cnt.value]


def wasserstein_ambiguous(u_amb: torch.Tensor, v_amb: torch.Tensor, tables: torch.Tensor,
                          nu_total: int, nv_total: int) -> float:
    """Exact integral over the ambiguous bins from all their values (any order)."""
    lib = _lib.load()
    u_amb, v_amb = _flat_f32(u_amb, "u_amb"), _flat_f32(v_amb, "v_amb")
    out = C.c_double()
    dev = tables.device
    with torch.cuda.device(dev):
        wsb = int(lib.uq_wasserstein_workspace_bytes(max(u_amb.numel(), 1), max(v_amb.numel(), 1)))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(lib.uq_wasserstein_ambiguous(
            u_amb.data_ptr() if u_amb.numel() else None, u_amb.numel(),
            v_amb.data_ptr() if v_amb.numel() else None, v_amb.numel(), tables.data_ptr(),
            int(nu_total), int(nv_total), C.byref(out), ws.data_ptr(), wsb, _stream_ptr(dev)))
    return float(out.value)


def wasserstein_1d_range(u: torch.Tensor, v: torch.Tensor, u_below: int, v_below: int,
                         nu_total: int, nv_total: int) -> Tuple[float, float, float]:
    """(partial integral, first merged value, last merged value) of one value range."""
    lib = _lib.load()
    u, v = _flat_f32(u, "u"), _flat_f32(v, "v")
    out = (C.c_double * 3)()
    with torch.cuda.device(u.device):
        wsb = int(lib.uq_wasserstein_workspace_bytes(max(u.numel(), 1), max(v.numel(), 1)))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=u.device)
        _lib.check(lib.uq_wasserstein_1d_range(
            u.data_ptr() if u.numel() else None, u.numel(), v.data_ptr() if v.numel() else None,
            v.numel(), int(u_below), int(v_below), int(nu_total), int(nv_total), out,
            ws.data_ptr(), wsb, _stream_ptr(u.device)))
    return float(out[0]), float(out[1]), float(out[2])


# ------------------------------------------------------------------------------------------------
# torch custom-op registration ("registered once so Python sees torch.ops.<ns>.*").
# The schemas carry tensors and plain numbers only; each implementation is registered for the
# CUDA dispatch key alone and ends in one call through the C ABI, so a CPU tensor reaching the
# dispatcher raises (NotImplementedError: no CPU kernel) instead of falling back.
# ------------------------------------------------------------------------------------------------

def _as_i64(v: int) -> int:
    v = int(v) & 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def _run_forward(lib, handle, x: torch.Tensor, a: "_lib.ForwardArgs", out0: torch.Tensor,
                 out1: torch.Tensor) -> None:
    n, dev = x.shape[0], x.device
    with _on_device(dev):
        wsb = int(lib.uq_forward_workspace_bytes(handle, n, C.byref(a)))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
        _lib.check(lib.uq_forward(handle, x.data_ptr(), n, C.byref(a), out0.data_ptr(),
                                  out1.data_ptr(), ws.data_ptr(), wsb, None, _stream_ptr(dev)))


def _op_uq_forward(handle: int, x: torch.Tensor, mode: int, precision: int, output: int,
                   member_begin: int, member_count: int, total_members: int,
                   dropout_active: bool, dropout_p: float, seed: int, offset: int,
                   masks: Optional[torch.Tensor], anchors: Optional[torch.Tensor], d_out: int,
                   targets: Optional[torch.Tensor] = None,
                   score_floor: Optional[torch.Tensor] = None, row_base: int = 0
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    a = _lib.ForwardArgs()
    a.mode, a.precision, a.output = mode, precision, output
    a.row_base = row_base
    a.member_begin, a.member_count, a.total_members = member_begin, member_count, total_members
    a.dropout_active = 1 if dropout_active else 0
    a.dropout_p = dropout_p
    a.philox_seed = seed & 0xFFFFFFFFFFFFFFFF
    a.philox_offset = offset & 0xFFFFFFFFFFFFFFFF
    a.masks = masks.data_ptr() if masks is not None else None
    a.anchors = anchors.data_ptr() if anchors is not None else None
    a.anchor_targets = targets.data_ptr() if targets is not None else None
    a.score_floor = score_floor.data_ptr() if score_floor is not None else None
    n, dev = x.shape[0], x.device
    out0 = torch.empty((n, d_out), dtype=torch.float32, device=dev)
    out1 = torch.empty((n, d_out), dtype=torch.float32, device=dev)
    _run_forward(lib, C.c_void_p(handle), x, a, out0, out1)
    return out0, out1


def _op_moments_merge(means: torch.Tensor, m2s: torch.Tensor, counts: Sequence[float]
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    s = means.shape[0]
    length = means[0].numel()
    out_mean = torch.empty(means.shape[1:], dtype=torch.float32, device=means.device)
    out_std = torch.empty_like(out_mean)
    cnt = (C.c_double * s)(*[float(c) for c in counts])
    with torch.cuda.device(means.device):
        _lib.check(lib.uq_moments_merge(means.data_ptr(), m2s.data_ptr(), cnt, s, length,
                                        out_mean.data_ptr(), out_std.data_ptr(),
                                        _stream_ptr(means.device)))
    return out_mean, out_std


def _op_wasserstein_1d(u: torch.Tensor, v: torch.Tensor, method: int = 0) -> float:
    lib = _lib.load()
    out = C.c_double()
    with _on_device(u.device):
        wsb = int(lib.uq_wasserstein_workspace_bytes(u.numel(), v.numel()))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=u.device)
        _lib.check(lib.uq_wasserstein_1d_ex(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(),
                                            method, C.byref(out), None, ws.data_ptr(), wsb,
                                            _stream_ptr(u.device)))
    return float(out.value)


def _op_kde_jsd(u: torch.Tensor, v: torch.Tensor, num_points: int, method: int = 0) -> float:
    lib = _lib.load()
    out = C.c_double()
    with _on_device(u.device):
        wsb = int(lib.uq_kde_jsd_workspace_bytes(u.numel(), v.numel(), num_points))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=u.device)
        _lib.check(lib.uq_kde_jsd_ex(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(), num_points,
                                     method, C.byref(out), None, ws.data_ptr(), wsb,
                                     _stream_ptr(u.device)))
    return float(out.value)


OP_NAMESPACE = "nnueehcs_b200"
OP_SCHEMAS = {
    "uq_forward": "(int handle, Tensor x, int mode, int precision, int output, int member_begin, "
                  "int member_count, int total_members, bool dropout_active, float dropout_p, "
                  "int seed, int offset, Tensor? masks, Tensor? anchors, int d_out, "
                  "Tensor? targets=None, Tensor? score_floor=None, int row_base=0) "
                  "-> (Tensor, Tensor)",
    "moments_merge": "(Tensor means, Tensor m2s, float[] counts) -> (Tensor, Tensor)",
    "wasserstein_1d": "(Tensor u, Tensor v, int method=0) -> float",
    "kde_jsd": "(Tensor u, Tensor v, int num_points, int method=0) -> float",
}
_OP_IMPLS = {"uq_forward": _op_uq_forward, "moments_merge": _op_moments_merge,
             "wasserstein_1d": _op_wasserstein_1d, "kde_jsd": _op_kde_jsd}
_op_library = torch.library.Library(OP_NAMESPACE, "DEF")
for _name, _schema in OP_SCHEMAS.items():
    _op_library.define(_name + _schema)
    _op_library.impl(_name, _OP_IMPLS[_name], "CUDA")
