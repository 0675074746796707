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


# ----------------------------------------------------------------------------------------------
# Score consumers: closed forms on the two SORTED score arrays of
# what the reference computes with Python loops / sklearn / torch.quantile.  Pinned by
# tests/golden/score_metrics.npz (outputs of the reference's own classes).
# ----------------------------------------------------------------------------------------------

def _min_count(target: float, denom: int) -> int:
    """Smallest integer m >= 0 with  m / denom >= target  in Python float arithmetic (the
    comparison ``tpr >= self.target_tpr`` of nnueehcs/evaluation.py:577)."""
    m = int(np.ceil(target * denom))
    while m > 0 and (m - 1) / denom >= target:
        m -= 1
    while m / denom < target:
        m += 1
    return m


def tnr_at_tpr(id_scores, ood_scores, target_tpr: float, reversed_: bool = False) -> float:
    """``TNRatTPX._evaluate_scores`` (nnueehcs/evaluation.py:538-580) without the loop over every
    unique score.  tnr is non-decreasing and tpr non-increasing in the threshold, so the best tnr
    is the one at the LARGEST threshold (a score value) whose tpr still meets the target; that
    threshold is the largest score below c = the m-th largest positive, and every negative below
    c is <= it.  Keeps the reference's quirks: in reversed mode tp counts ID scores but is divided
    by n_ood, and tn counts OOD scores but is divided by n_id."""
    a = np.sort(np.asarray(id_scores, dtype=np.float32).ravel())
    b = np.sort(np.asarray(ood_scores, dtype=np.float32).ravel())
    n_id, n_ood = a.size, b.size
    if reversed_:
        if a[0] > b[-1]:
            return 1.0
        pos, neg = a, b
    else:
        if a[-1] < b[0]:
            return 1.0
        pos, neg = b, a
    m = _min_count(target_tpr, n_ood)          # tp / n_ood >= target  <=>  tp >= m
    if m > pos.size:
        return 0.0
    if m == 0:                                  # every threshold qualifies: largest = max(all)
        return neg.size / n_id
    c = pos[pos.size - m]
    if not (min(a[0], b[0]) < c):               # no score below c: no threshold qualifies
        return 0.0
    return int(np.searchsorted(neg, c, side="left")) / n_id


def auroc(id_scores, ood_scores) -> float:
    """``AUROC._evaluate_scores`` (nnueehcs/evaluation.py:614-624): sklearn ``roc_auc_score`` with
    OOD as the positive class == Mann-Whitney U / (n_id n_ood), ties counted one half."""
    a = np.sort(np.asarray(id_scores, dtype=np.float32).ravel())
    b = np.asarray(ood_scores, dtype=np.float32).ravel()
    lt = np.searchsorted(a, b, side="left").astype(np.int64)
    le = np.searchsorted(a, b, side="right").astype(np.int64)
    return float((lt + le).sum()) / (2.0 * a.size * b.size)


def torch_quantile_f32(sorted_vals: np.ndarray, q: float) -> np.float32:
    """``torch.quantile(x, q)`` for a float32 vector, restated: float32 rank, float32 lerp with
    ATen's two-sided formula."""
    n = sorted_vals.size
    rank = np.float32(q) * np.float32(n - 1)
    below = np.floor(rank)
    above = np.ceil(rank)
    w = np.float32(rank - below)
    lo, hi = sorted_vals[int(below)], sorted_vals[int(above)]
    diff = np.float32(hi - lo)
    if w < np.float32(0.5):
        return np.float32(lo + np.float32(w * diff))
    return np.float32(hi - np.float32(diff * np.float32(np.float32(1.0) - w)))


def percentile_classifier(id_scores, ood_scores, percentile: float, reversed_: bool = False):
    """``PercentileBasedClassifier`` (evaluation.py:637-662) over
    ``PercentileBasedIdOodClassifier._evaluate_scores`` (classification.py:103-143): threshold =
    ``torch.quantile(id, percentile)`` (scores negated first when reversed), then four counts.
    Returns (sensitivity, specificity, fpr, fnr)."""
    a = np.asarray(id_scores, dtype=np.float32).ravel()
    b = np.asarray(ood_scores, dtype=np.float32).ravel()
    if reversed_:
        a, b = -a, -b
    a, b = np.sort(a), np.sort(b)
    thr = torch_quantile_f32(a, percentile)
    id_above = a.size - int(np.searchsorted(a, thr, side="right"))
    ood_above = b.size - int(np.searchsorted(b, thr, side="right"))
    id_below, ood_below = a.size - id_above, b.size - ood_above
    def ratio(x, y):
        return float(x) / (x + y) if (x + y) else 0.0
    return (ratio(ood_above, ood_below), ratio(id_below, id_above), ratio(id_above, id_below),
            ratio(ood_below, ood_above))


def score_summaries(id_scores, percentile_q: float):
    """(mean_score, max_score, percentile_score) as the reference computes them with numpy on the
    float32 ID scores (evaluation.py:303, :323, :365)."""
    a = np.asarray(id_scores, dtype=np.float32).ravel()
    return float(np.mean(a)), float(np.max(a)), float(np.percentile(a, percentile_q))


# ----------------------------------------------------------------------------------------------
# Bin-moment form of the same Wasserstein integral (csrc/wasserstein.cu, "binned" method).
# Not part of the reference: a numpy stand-in that the tests check the device tables and the
# device result against, itself checked against wasserstein_1d above (scipy's arithmetic).
#
# The real line is cut at the 16384 key-bin boundaries t_b.  Inside one bin every float32 is
# t_b + k * ulp_b with k = low 18 key bits, so  integral_bin F_u = ((C_b + c_b) w_b - K_b ulp_b)/n
# from the bin's count c_b, the count below C_b and the integer sum K_b = sum k.  Where
# D = F_u - F_v provably keeps one sign across the bin, |integral D| is the bin's contribution;
# the other ("ambiguous") bins are integrated exactly from their sorted values.
# ----------------------------------------------------------------------------------------------

def _keys(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32) + np.float32(0.0)  # -0.0 -> +0.0
    b = x.view(np.uint32)
    return np.where(b >> 31, ~b, b ^ np.uint32(0x80000000)).astype(np.uint32)


def bin_edges() -> tuple:
    """(t[KEY_BINS + 1], ulp[KEY_BINS]) as float64: lower edge and float spacing of every bin."""
    b = np.arange(KEY_BINS + 1, dtype=np.uint64)
    key = b << np.uint64(18)
    pos = key >= np.uint64(0x80000000)
    bits = np.where(pos, key - np.uint64(0x80000000), (~key) & np.uint64(0xFFFFFFFF) & np.uint64(0x7FFFFFFF))
    e = (bits >> np.uint64(23)).astype(np.int64)          # 0 .. 256 (256 only for b = KEY_BINS)
    m = (bits & np.uint64(0x7FFFFF)).astype(np.float64)
    mag = np.where(e == 0, np.ldexp(m, -149), np.ldexp(m + 8388608.0, e - 150))
    t = np.where(pos, mag, -mag)
    ulp = np.ldexp(1.0, np.maximum(e[:-1], 1) - 150)
    return t, ulp


def bin_moments(x: np.ndarray) -> tuple:
    """(count[KEY_BINS], ksum[KEY_BINS]) int64: values per bin and sum of their low 18 key bits."""
    k = _keys(np.asarray(x).ravel())
    b = (k >> np.uint32(18)).astype(np.int64)
    low = (k & np.uint32(0x3FFFF)).astype(np.int64)
    cnt = np.bincount(b, minlength=KEY_BINS).astype(np.int64)
    ks = np.bincount(b, weights=low.astype(np.float64), minlength=KEY_BINS)
    return cnt, np.rint(ks).astype(np.int64)


def bin_resolve(cu, ku, cv, kv) -> dict:
    """Per-bin contribution of the sign-definite bins, the ambiguity flags and the rank offsets
    the exact pass over the ambiguous bins needs (python ints: products reach 2^62)."""
    nu, nv = int(cu.sum()), int(cv.sum())
    t, ulp = bin_edges()
    Cu = np.concatenate([[0], np.cumsum(cu)]).astype(object)
    Cv = np.concatenate([[0], np.cumsum(cv)]).astype(object)
    flags = np.zeros(KEY_BINS, dtype=np.uint8)
    resolved = 0.0
    for b in np.nonzero((cu > 0) | (cv > 0) | (Cu[:-1] * nv != Cv[:-1] * nu))[0]:
        b = int(b)
        d_ge0 = Cu[b] * nv - (Cv[b] + int(cv[b])) * nu >= 0
        d_le0 = (Cu[b] + int(cu[b])) * nv - Cv[b] * nu <= 0
        if d_ge0 or d_le0:
            w = t[b + 1] - t[b]
            a = (float(Cu[b] + int(cu[b])) * w - float(ku[b]) * ulp[b]) / float(nu)
            c = (float(Cv[b] + int(cv[b])) * w - float(kv[b]) * ulp[b]) / float(nv)
            resolved += abs(a - c)
        else:
            flags[b] = 1
    amb_u = np.concatenate([[0], np.cumsum(cu * flags)])
    amb_v = np.concatenate([[0], np.cumsum(cv * flags)])
    skip_u = np.asarray(Cu, dtype=np.int64) - amb_u
    skip_v = np.asarray(Cv, dtype=np.int64) - amb_v
    return {"resolved": resolved, "flags": flags, "skip_u": skip_u, "skip_v": skip_v,
            "amb_u": int(amb_u[-1]), "amb_v": int(amb_v[-1]), "t": t}


def wasserstein_ambiguous(u_amb, v_amb, cu, ku, cv, kv) -> float:
    """Exact integral over the ambiguous bins from all their values (bin tables of the whole
    samples give the ranks below each bin)."""
    r = bin_resolve(cu, ku, cv, kv)
    t, flags = r["t"], r["flags"]
    nu, nv = float(cu.sum()), float(cv.sum())
    u = np.asarray(u_amb, dtype=np.float32).ravel()
    v = np.asarray(v_amb, dtype=np.float32).ravel()
    bu = (_keys(u) >> np.uint32(18)).astype(np.int64)
    bv = (_keys(v) >> np.uint32(18)).astype(np.int64)
    total = 0.0
    for b in np.nonzero(flags)[0]:
        xu = np.sort(u[bu == b].astype(np.float64))
        xv = np.sort(v[bv == b].astype(np.float64))
        allv = np.sort(np.concatenate([xu, xv]), kind="mergesort")
        pts = np.concatenate([[t[b]], allv, [t[b + 1]]])
        ru = np.concatenate([[0], np.searchsorted(xu, allv, side="right")])
        rv = np.concatenate([[0], np.searchsorted(xv, allv, side="right")])
        cb, db = int(cu[:b].sum()), int(cv[:b].sum())
        d = (cb + ru) / nu - (db + rv) / nv
        total += float(np.sum(np.abs(d) * np.diff(pts)))
    return total


def wasserstein_1d_binned(u, v) -> float:
    u = np.asarray(u, dtype=np.float32).ravel()
    v = np.asarray(v, dtype=np.float32).ravel()
    cu, ku = bin_moments(u)
    cv, kv = bin_moments(v)
    r = bin_resolve(cu, ku, cv, kv)
    total = r["resolved"]
    if r["amb_u"] + r["amb_v"]:
        flags = r["flags"].astype(bool)
        total += wasserstein_ambiguous(u[flags[key_bin(u + np.float32(0.0))]],
                                       v[flags[key_bin(v + np.float32(0.0))]], cu, ku, cv, kv)
    return float(total)


def _np_bin_moments(x, tables, row):
    import torch
    if x.numel() == 0:
        return
    c, k = bin_moments(x.detach().cpu().numpy())
    tables[row] += torch.from_numpy(c)
    tables[row + 1] += torch.from_numpy(k)


def _np_from_bins(tables, nu_total, nv_total):
    import torch
    t = tables.cpu().numpy()
    assert int(t[0].sum()) == nu_total and int(t[2].sum()) == nv_total
    r = bin_resolve(t[0], t[1], t[2], t[3])
    nonfinite = int(t[0][:32].sum() + t[0][-32:].sum() + t[2][:32].sum() + t[2][-32:].sum())
    return {"resolved": float(r["resolved"]), "amb_u": r["amb_u"], "amb_v": r["amb_v"],
            "nonfinite": nonfinite, "flags": torch.from_numpy(r["flags"])}


def _np_compact_flagged(x, flags):
    import torch
    a = x.detach().cpu().numpy().ravel()
    keep = flags.cpu().numpy().astype(bool)[key_bin(a + np.float32(0.0))]
    return torch.from_numpy(a[keep].copy())


def _np_ambiguous(u_amb, v_amb, tables, nu_total, nv_total):
    t = tables.cpu().numpy()
    return wasserstein_ambiguous(u_amb.cpu().numpy(), v_amb.cpu().numpy(), t[0], t[1], t[2], t[3])


NumpyShardBackend.bin_moments = staticmethod(_np_bin_moments)
NumpyShardBackend.wasserstein_from_bins = staticmethod(_np_from_bins)
NumpyShardBackend.compact_flagged = staticmethod(_np_compact_flagged)
NumpyShardBackend.wasserstein_ambiguous = staticmethod(_np_ambiguous)


# ----------------------------------------------------------------------------------------------
# KDEMLPModel's input-density score (nnueehcs/models.py:191-222): sklearn KernelDensity restated.
# Pinned by tests/golden/kde_density.npz (outputs of the reference's own KDEMLPModel + sklearn).
# ----------------------------------------------------------------------------------------------

def scott_bandwidth_sklearn(m: int, d: int) -> float:
    """sklearn/neighbors/_kde.py, KernelDensity.fit: bandwidth_ = n_samples ** (-1 / (n_features + 4))."""
    return float(m) ** (-1.0 / (d + 4))


def silverman_bandwidth_sklearn(m: int, d: int) -> float:
    """sklearn/neighbors/_kde.py, KernelDensity.fit, bandwidth='silverman':
    bandwidth_ = (n_samples * (n_features + 2) / 4) ** (-1 / (n_features + 4))."""
    return (float(m) * (d + 2) / 4.0) ** (-1.0 / (d + 4))


def kde_neg_density(fit: np.ndarray, x: np.ndarray, bandwidth: float, block: int = 256) -> np.ndarray:
    """``-exp(KernelDensity(bandwidth=h, kernel='gaussian').fit(fit).score_samples(x))`` in float64:
    log-density = logsumexp(-|x - y|^2 / (2 h^2)) - log(m) - d log(h) - d/2 log(2 pi)."""
    fit = np.asarray(fit, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    m, d = fit.shape
    out = np.empty(x.shape[0])
    log_norm = -np.log(m) - d * np.log(bandwidth) - 0.5 * d * np.log(2.0 * np.pi)
    for i in range(0, x.shape[0], block):
        d2 = ((x[i:i + block, None, :] - fit[None, :, :]) ** 2).sum(-1)
        e = -0.5 * d2 / bandwidth ** 2
        mx = e.max(axis=1, keepdims=True)
        out[i:i + block] = (mx[:, 0] + np.log(np.exp(e - mx).sum(axis=1))) + log_norm
    return -np.exp(out)


# ----------------------------------------------------------------------------------------------
# Moment form of the KDE sums (csrc/kde_jsd.cu, "moments" method).  Not part of the reference: a
# numpy stand-in of the product's default path, itself checked against pdf_jsd above (scipy's
# arithmetic).  Bins of width h / 4; a value at offset eps * h from its bin centre c contributes
#   exp(-((g - c)/h - eps)^2 / 2) = K(z) exp(z eps - eps^2 / 2) = K(z) sum_m He_m(z) eps^m / m!
# (z = (g - c)/h, He_m the probabilists' Hermite polynomials), so the sum over a bin's values needs
# only the bin's count and sum eps^m, m = 1..5.
# ----------------------------------------------------------------------------------------------

KM_PER_H, KM_ORDER, KM_MAX_BINS = 4, 5, 8192


def kde_sums_moments(x: np.ndarray, lo: float, hi: float, h: float, num_points: int) -> np.ndarray:
    """Raw Gaussian kernel sums of ``x`` on ``linspace(lo, hi, num_points)`` via per-bin moments."""
    import math
    x = np.asarray(x, dtype=np.float64).ravel()
    w = h / KM_PER_H
    nb = int(np.floor((hi - lo) / w)) + 1
    if nb > KM_MAX_BINS:
        raise ValueError("the moment method needs (max - min) / (h / 4) <= 8192 bins")
    b = np.clip(((x - lo) / w).astype(np.int64), 0, nb - 1)
    eps = (x - (lo + (b + 0.5) * w)) / h
    mom = np.stack([np.bincount(b, weights=eps ** m, minlength=nb) / math.factorial(m)
                    for m in range(KM_ORDER + 1)])
    centres = lo + (np.arange(nb) + 0.5) * w
    grid = np.linspace(lo, hi, num_points)
    out = np.zeros(num_points)
    for j0 in range(0, num_points, 1024):
        z = (grid[j0:j0 + 1024, None] - centres[None, :]) / h
        he = [np.ones_like(z), z]
        for m in range(2, KM_ORDER + 1):
            he.append(z * he[-1] - (m - 1) * he[-2])
        series = sum(he[m] * mom[m][None, :] for m in range(KM_ORDER + 1))
        out[j0:j0 + 1024] = (np.where(np.abs(z) <= 9.0 + 0.5 / KM_PER_H, np.exp(-0.5 * z * z), 0.0)
                             * series).sum(axis=1)
    return out


def pdf_jsd_moments(dist1: np.ndarray, dist2: np.ndarray, num_points: int = 20000) -> float:
    a = np.asarray(dist1, dtype=np.float64).ravel()
    c = np.asarray(dist2, dtype=np.float64).ravel()
    lo, hi = float(min(a.min(), c.min())), float(max(a.max(), c.max()))
    return jensenshannon(kde_sums_moments(a, lo, hi, scott_bandwidth(a), num_points),
                         kde_sums_moments(c, lo, hi, scott_bandwidth(c), num_points))
