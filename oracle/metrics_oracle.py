"""numpy restatement of the two distribution metrics on the hot path.  TEST INFRASTRUCTURE ONLY.

The reference computes both through scipy (unpinned in its ``pyproject.toml:17``; scipy 1.18.1
in the authoring container):

* ``WassersteinEvaluation._evaluate_uncertainties`` -> ``scipy.stats.wasserstein_distance``
  (nnueehcs/evaluation.py:175-188).  Algorithm restated from
  ``scipy/stats/_stats_py.py:_cdf_distance`` (p = 1).
* ``JensenShannonEvaluation.pdf_jsd`` -> two ``scipy.stats.gaussian_kde`` evaluated on a shared
  20 000-point ``linspace`` and ``scipy.spatial.distance.jensenshannon``
  (nnueehcs/evaluation.py:268-276).  KDE restated from ``scipy/stats/_kde.py`` (Scott factor
  ``n**(-1/5)``, unbiased data variance, float64 evaluation).

Pinned against scipy itself in ``tests/golden/metrics_*.npz`` (``make_golden.py``) and, when
scipy is importable, directly in ``tests/test_oracle.py``.
"""
from __future__ import annotations

import numpy as np


def wasserstein_1d(u_values: np.ndarray, v_values: np.ndarray) -> float:
    """W1 = integral |U - V| between the empirical CDFs of two unweighted samples.

    Follows ``_cdf_distance(p=1, ...)``: ``_validate_distribution`` first converts the values to
    float64; then sort each sample, sort the concatenation, take successive differences, locate
    each merged value in both sorted samples with a right-sided search, and sum
    |cdf_u - cdf_v| * delta, everything in float64.
    """
    u = np.asarray(u_values, dtype=np.float64).ravel()
    v = np.asarray(v_values, dtype=np.float64).ravel()
    if u.size == 0 or v.size == 0:
        raise ValueError("Distribution can't be empty.")
    u_sorted = np.sort(u)
    v_sorted = np.sort(v)
    all_values = np.concatenate((u, v))
    all_values.sort(kind="mergesort")
    deltas = np.diff(all_values)
    u_cdf = u_sorted.searchsorted(all_values[:-1], "right") / u.size
    v_cdf = v_sorted.searchsorted(all_values[:-1], "right") / v.size
    return float(np.sum(np.multiply(np.abs(u_cdf - v_cdf), deltas)))


def scott_bandwidth(data: np.ndarray) -> float:
    """Kernel std h = sqrt(unbiased var) * n**(-1/5)  (``_kde.py`` scotts_factor, d = 1;
    ``_compute_covariance`` with uniform weights reduces to the unbiased variance)."""
    x = np.asarray(data, dtype=np.float64).ravel()
    n = x.size
    return float(np.sqrt(np.var(x, ddof=1)) * n ** (-1.0 / 5.0))


def gaussian_kde_pdf(data: np.ndarray, grid: np.ndarray, block: int = 2048) -> np.ndarray:
    """pdf(g) = 1/(n h sqrt(2 pi)) * sum_i exp(-((x_i - g)/h)^2 / 2), float64 throughout."""
    x = np.asarray(data, dtype=np.float64).ravel()
    g = np.asarray(grid, dtype=np.float64).ravel()
    h = scott_bandwidth(x)
    xs = x / h
    gs = g / h
    out = np.zeros_like(g)
    for s in range(0, xs.size, block):
        d = xs[s:s + block, None] - gs[None, :]
        out += np.exp(-0.5 * d * d).sum(axis=0)
    return out / (x.size * h * np.sqrt(2.0 * np.pi))


def rel_entr(p: np.ndarray, q: np.ndarray) -> np.ndarray:
    """scipy.special.rel_entr: p*log(p/q) for p>0,q>0; 0 for p==0,q>=0; inf otherwise."""
    out = np.full(p.shape, np.inf, dtype=np.float64)
    pos = (p > 0) & (q > 0)
    out[pos] = p[pos] * np.log(p[pos] / q[pos])
    out[(p == 0) & (q >= 0)] = 0.0
    return out


def jensenshannon(p: np.ndarray, q: np.ndarray) -> float:
    """``scipy.spatial.distance.jensenshannon`` (natural log): each vector is normalised to
    sum 1, m = (p+q)/2, distance = sqrt((KL(p||m) + KL(q||m)) / 2)."""
    p = np.asarray(p, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    p = p / p.sum()
    q = q / q.sum()
    m = (p + q) / 2.0
    js = rel_entr(p, m).sum() + rel_entr(q, m).sum()
    return float(np.sqrt(js / 2.0))


def pdf_jsd(dist1: np.ndarray, dist2: np.ndarray, num_points: int = 20000) -> float:
    """``JensenShannonEvaluation.pdf_jsd`` (nnueehcs/evaluation.py:268-276)."""
    d1 = np.asarray(dist1).ravel()
    d2 = np.asarray(dist2).ravel()
    lo = min(d1.min(), d2.min())
    hi = max(d1.max(), d2.max())
    x_range = np.linspace(lo, hi, num_points)
    return jensenshannon(gaussian_kde_pdf(d1, x_range), gaussian_kde_pdf(d2, x_range))


# ----------------------------------------------------------------------------------------------
# numpy stand-ins for the per-rank CUDA steps of the sharded metrics (TEST INFRASTRUCTURE):
# they let the world-size-2 gloo tests exercise nnueehcs_b200.distributed's host logic on CPU,
# and the GPU tests check each CUDA step against them.
# ----------------------------------------------------------------------------------------------

KEY_BINS = 16384


def key_bin(x: np.ndarray) -> np.ndarray:
    """Top 14 bits of the order-preserving radix key of float32 values (csrc/shard_metrics.cu)."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    k = np.where(b >> 31, ~b, b ^ np.uint32(0x80000000)).astype(np.uint32)
    return (k >> 18).astype(np.int64)


def wasserstein_1d_range(u, v, u_below, v_below, nu_total, nv_total):
    """Partial integral of |F_u - F_v| over one value range with global CDF offsets."""
    u = np.sort(np.asarray(u, dtype=np.float64).ravel())
    v = np.sort(np.asarray(v, dtype=np.float64).ravel())
    allv = np.sort(np.concatenate([u, v]), kind="mergesort")
    if allv.size == 0:
        return 0.0, 0.0, 0.0
    deltas = np.diff(allv)
    cu = (u_below + np.searchsorted(u, allv[:-1], side="right")) / nu_total
    cv = (v_below + np.searchsorted(v, allv[:-1], side="right")) / nv_total
    return float(np.sum(np.abs(cu - cv) * deltas)), float(allv[0]), float(allv[-1])


class NumpyShardBackend:
    """Same interface as ``nnueehcs_b200.distributed.CudaMetricBackend`` on CPU torch tensors."""

    @staticmethod
    def key_bins():
        return KEY_BINS

    @staticmethod
    def sample_stats(x):
        a = x.detach().cpu().numpy().astype(np.float64).ravel()
        mean = a.mean()
        return float(a.min()), float(a.max()), float(mean), float(((a - mean) ** 2).sum())

    @staticmethod
    def kde_grid_accumulate(x, lo, hi, bandwidth, grid):
        import torch
        a = x.detach().cpu().numpy().astype(np.float64).ravel()
        g = np.linspace(lo, hi, grid.numel())
        acc = np.zeros(g.size)
        for i in range(0, a.size, 4096):
            z = (a[i:i + 4096, None] - g[None, :]) / bandwidth
            acc += np.exp(-0.5 * z * z).sum(0)
        grid += torch.from_numpy(acc).to(grid.device)

    @staticmethod
    def jsd_from_grids(grids):
        p, q = grids[0].cpu().numpy(), grids[1].cpu().numpy()
        p, q = p / p.sum(), q / q.sum()
        m = (p + q) / 2.0
        with np.errstate(divide="ignore", invalid="ignore"):
            left = np.where((p > 0) & (m > 0), p * np.log(p / m), 0.0)
            right = np.where((q > 0) & (m > 0), q * np.log(q / m), 0.0)
        return float(np.sqrt((left.sum() + right.sum()) / 2.0))

    @staticmethod
    def key_histogram(x):
        import torch
        a = x.detach().cpu().numpy().ravel()
        return torch.from_numpy(np.bincount(key_bin(a), minlength=KEY_BINS).astype(np.int64))

    @staticmethod
    def partition_by_bin(x, bin_to_part, part_counts):
        import torch
        a = x.detach().cpu().numpy().ravel()
        dst = bin_to_part.cpu().numpy()[key_bin(a)]
        order = np.argsort(dst, kind="stable")
        assert [int((dst == p).sum()) for p in range(len(part_counts))] == list(part_counts)
        return torch.from_numpy(a[order].copy())

    @staticmethod
    def wasserstein_1d_range(u, v, u_below, v_below, nu_total, nv_total):
        return wasserstein_1d_range(u.cpu().numpy(), v.cpu().numpy(), u_below, v_below, nu_total,
                                    nv_total)
