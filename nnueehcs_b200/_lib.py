"""ctypes binding of ``include/nnueehcs_b200.h``.

The product path has no CPU fallback: if the native library is not built this module raises,
and every op raises with it.  Build with ``python -m nnueehcs_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

ABI_VERSION = 3
UQ_OK, UQ_ERR_INVALID, UQ_ERR_CUDA, UQ_ERR_UNSUPPORTED, UQ_ERR_WORKSPACE = 0, 1, 2, 3, 4
MODE_ENSEMBLE, MODE_MC_DROPOUT, MODE_DELTA_UQ, MODE_PAGER = 0, 1, 2, 3
PREC_FP32, PREC_BF16, PREC_FP32_FFMA = 0, 1, 2
OUT_MEAN_STD, OUT_MOMENTS = 0, 1
MODEL_ANCHOR_FIRST = 1
WASSERSTEIN_AUTO, WASSERSTEIN_SORT, WASSERSTEIN_BINNED = 0, 1, 2
KDE_AUTO, KDE_WINDOW, KDE_MOMENTS = 0, 1, 2
METRIC_RECORD_BYTES = 256

# every symbol include/nnueehcs_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "uq_abi_version", "uq_last_error", "uq_launch_count", "uq_launch_count_reset",
    "uq_kde_jsd_phase_us",
    "uq_model_create", "uq_model_create_ex", "uq_model_destroy", "uq_model_supports_bf16", "uq_model_supports_fp32_tc",
    "uq_forward_workspace_bytes", "uq_forward", "uq_forward_host", "uq_moments_merge",
    "uq_moments_merge_ex",
    "uq_philox_keep_masks", "uq_wasserstein_workspace_bytes", "uq_wasserstein_1d",
    "uq_wasserstein_1d_ex",
    "uq_kde_jsd_workspace_bytes", "uq_kde_jsd", "uq_kde_jsd_ex",
    "uq_sample_stats_workspace_bytes", "uq_sample_stats", "uq_kde_grid_workspace_bytes",
    "uq_kde_grid_accumulate", "uq_jsd_from_grids", "uq_key_bins", "uq_key_histogram",
    "uq_partition_by_bin", "uq_wasserstein_1d_range",
    "uq_score_metrics_workspace_bytes", "uq_score_metrics",
    "uq_kde_scott_bandwidth", "uq_kde_silverman_bandwidth", "uq_kde_density_workspace_bytes",
    "uq_kde_density",
    "uq_sort_workspace_bytes", "uq_sort_f32",
    "uq_wasserstein_1d_enqueue", "uq_wasserstein_1d_finish", "uq_kde_jsd_enqueue",
    "uq_kde_jsd_finish",
    "uq_bin_moments", "uq_wasserstein_from_bins", "uq_compact_flagged", "uq_wasserstein_ambiguous",
)


class LayerDesc(C.Structure):
    _fields_ = [
        ("in_features", C.c_int32),
        ("out_features", C.c_int32),
        ("weight", C.c_void_p),
        ("bias", C.c_void_p),
        ("bn_weight", C.c_void_p),
        ("bn_bias", C.c_void_p),
        ("bn_mean", C.c_void_p),
        ("bn_var", C.c_void_p),
        ("bn_eps", C.c_float),
        ("relu", C.c_int32),
        ("dropout", C.c_int32),
    ]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("precision", C.c_int32),
        ("output", C.c_int32),
        ("member_begin", C.c_int32),
        ("member_count", C.c_int32),
        ("total_members", C.c_int32),
        ("dropout_active", C.c_int32),
        ("row_base", C.c_int32),
        ("dropout_p", C.c_double),
        ("philox_seed", C.c_uint64),
        ("philox_offset", C.c_uint64),
        ("masks", C.c_void_p),
        ("anchors", C.c_void_p),
        ("anchor_targets", C.c_void_p),
        ("score_floor", C.c_void_p),
    ]


class ScoreRequest(C.Structure):
    _fields_ = [("percentile_q", C.c_double), ("target_tpr", C.c_double),
                ("classifier_percentile", C.c_double), ("tnr_reversed", C.c_int32),
                ("classifier_reversed", C.c_int32)]


class ScoreResult(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("mean_score", "max_score", "percentile_score", "auroc",
                                          "tnr_at_tpr", "sensitivity", "specificity", "fpr", "fnr")]


_lib = None


def load() -> C.CDLL:
    """Load the native library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("NNUEEHCS_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError(
            f"nnueehcs_b200 native library not found at {path}; run "
            "`python -m nnueehcs_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64, u64, sz, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_size_t, C.c_double
    lib.uq_abi_version.restype = C.c_int
    lib.uq_last_error.restype = C.c_char_p
    lib.uq_launch_count.restype = u64
    lib.uq_launch_count_reset.restype = None
    lib.uq_model_create.argtypes = [C.POINTER(vp), i32, i32, C.POINTER(LayerDesc), vp]
    lib.uq_model_create_ex.argtypes = [C.POINTER(vp), i32, i32, C.POINTER(LayerDesc), i32, vp]
    lib.uq_model_create_ex.restype = C.c_int
    lib.uq_model_destroy.argtypes = [vp]
    lib.uq_model_supports_bf16.argtypes = [vp]
    lib.uq_model_supports_fp32_tc.argtypes = [vp]
    lib.uq_model_supports_fp32_tc.restype = C.c_int
    lib.uq_forward_workspace_bytes.argtypes = [vp, i64, C.POINTER(ForwardArgs)]
    lib.uq_forward_workspace_bytes.restype = sz
    lib.uq_forward.argtypes = [vp, vp, i64, C.POINTER(ForwardArgs), vp, vp, vp, sz,
                               C.POINTER(dbl), vp]
    lib.uq_forward_host.argtypes = [vp, vp, i64, C.POINTER(ForwardArgs), vp, vp, vp]
    lib.uq_moments_merge.argtypes = [vp, vp, C.POINTER(dbl), i32, i64, vp, vp, vp]
    lib.uq_moments_merge_ex.argtypes = [vp, vp, i64, C.POINTER(dbl), i32, i64, vp, vp, i32, vp]
    lib.uq_moments_merge_ex.restype = C.c_int
    lib.uq_philox_keep_masks.argtypes = [vp, i64, i32, i32, i32, dbl, u64, u64, vp]
    lib.uq_wasserstein_workspace_bytes.argtypes = [i64, i64]
    lib.uq_wasserstein_workspace_bytes.restype = sz
    lib.uq_wasserstein_1d.argtypes = [vp, i64, vp, i64, C.POINTER(dbl), vp, sz, vp]
    lib.uq_wasserstein_1d_ex.argtypes = [vp, i64, vp, i64, i32, C.POINTER(dbl), C.POINTER(i64), vp,
                                         sz, vp]
    lib.uq_wasserstein_1d_ex.restype = C.c_int
    lib.uq_kde_jsd_workspace_bytes.argtypes = [i64, i64, i32]
    lib.uq_kde_jsd_workspace_bytes.restype = sz
    lib.uq_kde_jsd.argtypes = [vp, i64, vp, i64, i32, C.POINTER(dbl), vp, sz, vp]
    lib.uq_kde_jsd_ex.argtypes = [vp, i64, vp, i64, i32, i32, C.POINTER(dbl), C.POINTER(i32), vp, sz,
                                  vp]
    lib.uq_kde_jsd_ex.restype = C.c_int
    lib.uq_sample_stats_workspace_bytes.restype = sz
    lib.uq_sample_stats.argtypes = [vp, i64, C.POINTER(dbl), vp, sz, vp]
    lib.uq_kde_grid_workspace_bytes.argtypes = [i64]
    lib.uq_kde_grid_workspace_bytes.restype = sz
    lib.uq_kde_grid_accumulate.argtypes = [vp, i64, dbl, dbl, dbl, i32, vp, vp, sz, vp]
    lib.uq_jsd_from_grids.argtypes = [vp, i32, C.POINTER(dbl), vp]
    lib.uq_key_bins.restype = i32
    lib.uq_key_histogram.argtypes = [vp, i64, vp, vp]
    lib.uq_partition_by_bin.argtypes = [vp, i64, vp, i32, vp, vp, vp]
    lib.uq_wasserstein_1d_range.argtypes = [vp, i64, vp, i64, i64, i64, i64, i64, C.POINTER(dbl),
                                            vp, sz, vp]
    lib.uq_kde_scott_bandwidth.argtypes = [i64, i32]
    lib.uq_kde_scott_bandwidth.restype = dbl
    lib.uq_kde_silverman_bandwidth.argtypes = [i64, i32]
    lib.uq_kde_silverman_bandwidth.restype = dbl
    lib.uq_kde_density_workspace_bytes.argtypes = [i64, i64]
    lib.uq_kde_density_workspace_bytes.restype = sz
    lib.uq_kde_density.argtypes = [vp, i64, vp, i64, i32, dbl, vp, vp, sz, vp]
    lib.uq_kde_density.restype = C.c_int
    lib.uq_bin_moments.argtypes = [vp, i64, vp, vp, vp]
    lib.uq_wasserstein_from_bins.argtypes = [vp, i64, i64, vp, C.POINTER(dbl), vp, sz, vp]
    lib.uq_compact_flagged.argtypes = [vp, i64, vp, vp, C.POINTER(i64), vp, sz, vp]
    lib.uq_wasserstein_ambiguous.argtypes = [vp, i64, vp, i64, vp, i64, i64, C.POINTER(dbl), vp,
                                             sz, vp]
    for name in ("uq_bin_moments", "uq_wasserstein_from_bins", "uq_compact_flagged",
                 "uq_wasserstein_ambiguous"):
        getattr(lib, name).restype = C.c_int
    lib.uq_wasserstein_1d_enqueue.argtypes = [vp, i64, vp, i64, vp, vp, sz, vp]
    lib.uq_wasserstein_1d_finish.argtypes = [vp, i64, vp, i64, vp, C.POINTER(dbl), C.POINTER(i64),
                                             vp, sz, vp]
    lib.uq_kde_jsd_enqueue.argtypes = [vp, i64, vp, i64, i32, vp, vp, sz, vp]
    lib.uq_kde_jsd_finish.argtypes = [vp, i64, vp, i64, i32, vp, C.POINTER(dbl), C.POINTER(i32),
                                      vp, sz, vp]
    for name in ("uq_wasserstein_1d_enqueue", "uq_wasserstein_1d_finish", "uq_kde_jsd_enqueue",
                 "uq_kde_jsd_finish"):
        getattr(lib, name).restype = C.c_int
    lib.uq_sort_workspace_bytes.argtypes = [i64]
    lib.uq_sort_workspace_bytes.restype = sz
    lib.uq_sort_f32.argtypes = [vp, i64, vp, vp, sz, vp]
    lib.uq_sort_f32.restype = C.c_int
    lib.uq_score_metrics_workspace_bytes.argtypes = [i64, i64]
    lib.uq_score_metrics_workspace_bytes.restype = sz
    lib.uq_score_metrics.argtypes = [vp, i64, vp, i64, C.POINTER(ScoreRequest),
                                     C.POINTER(ScoreResult), vp, sz, vp]
    lib.uq_score_metrics.restype = C.c_int
    for name in ("uq_sample_stats", "uq_kde_grid_accumulate", "uq_jsd_from_grids",
                 "uq_key_histogram", "uq_partition_by_bin", "uq_wasserstein_1d_range"):
        getattr(lib, name).restype = C.c_int
    for name in ("uq_model_create", "uq_model_create_ex", "uq_model_destroy", "uq_model_supports_bf16", "uq_forward",
                 "uq_forward_host", "uq_moments_merge", "uq_philox_keep_masks",
                 "uq_wasserstein_1d", "uq_kde_jsd"):
        getattr(lib, name).restype = C.c_int
    if lib.uq_abi_version() != ABI_VERSION:
        raise RuntimeError("nnueehcs_b200: ABI version mismatch between _lib.py and the library")
    _lib = lib
    return lib


def check(status: int) -> None:
    """Map a C status to the exception kinds the reference's callers handle."""
    if status == UQ_OK:
        return
    msg = load().uq_last_error().decode("utf-8", "replace")
    if status in (UQ_ERR_INVALID, UQ_ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)
