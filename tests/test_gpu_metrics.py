"""Parity of the CUDA metric kernels with scipy (via goldens and the numpy oracle).  B200 only.

Wasserstein: the kernel performs the same float64 arithmetic as scipy on the same sorted values,
only the summation order differs -> relative 1e-12.
KDE-JS: scipy evaluates every kernel term in float64; the kernel evaluates them in float32 on
window-relative coordinates with MUFU.EX2 and truncates beyond 9 bandwidths -> stated tolerance
relative 2e-5 on the distance (measured values are printed).
"""
import numpy as np
import pytest
import torch

from nnueehcs_b200 import evaluation, ops
from oracle import metrics_oracle
from tests.util import load_golden

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
JSD_RTOL = 2e-5


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("name", ["metrics_small.npz", "metrics_medium.npz"])
def test_wasserstein_matches_scipy_golden(name):
    g = load_golden(name)
    w = ops.wasserstein_1d(_dev(g["id"]), _dev(g["ood"]))
    assert w == pytest.approx(float(g["wasserstein"]), rel=1e-12)
    assert ops.wasserstein_1d(_dev(g["id"]), _dev(g["id"])) == 0.0


@pytest.mark.parametrize("nu,nv", [(1, 1), (1, 7), (5, 3), (4096, 4097), (10000, 333),
                                   (70001, 130003)])
def test_wasserstein_sizes_ties_and_negatives(nu, nv):
    rng = np.random.default_rng(nu * 7 + nv)
    u = rng.normal(0.0, 1.0, nu).astype(np.float32)
    v = (rng.normal(0.2, 2.0, nv)).astype(np.float32)
    if nu > 10:
        u[: nu // 3] = np.round(u[: nu // 3], 1)  # heavy ties, including -0.0 / +0.0
        v[: nv // 3] = np.round(v[: nv // 3], 1)
    ref = metrics_oracle.wasserstein_1d(u, v)
    got = ops.wasserstein_1d(_dev(u), _dev(v))
    assert got == pytest.approx(ref, rel=1e-11, abs=1e-15)
    assert ops.wasserstein_1d(_dev(v), _dev(u)) == pytest.approx(ref, rel=1e-11, abs=1e-15)


def test_wasserstein_properties_and_reference_edge_cases():
    # reference tests/test_evaluation.py:189-208, :296-314
    a = torch.linspace(0, 1, 1000, device=DEV)
    assert ops.wasserstein_1d(a, a) == 0.0
    assert ops.wasserstein_1d(a, a + 5.0) == pytest.approx(5.0, rel=1e-6)
    ext = torch.tensor([1e-10, 1e10, 1.0, 3.0], device=DEV)
    assert np.isfinite(ops.wasserstein_1d(ext, ext.flip(0) * 2))
    with pytest.raises(ValueError, match="can't be empty"):
        ops.wasserstein_1d(a[:0], a)
    # input tensors are left untouched (the kernel sorts copies)
    b = torch.rand(5000, device=DEV)
    keep = b.clone()
    ops.wasserstein_1d(b, a)
    assert torch.equal(b, keep)


def _binned_cases():
    rng = np.random.default_rng(11)
    g2 = rng.gamma(2.0, 0.05, 300001).astype(np.float32)
    g3 = rng.gamma(3.0, 0.08, 250000).astype(np.float32)
    nrm = rng.normal(0.0, 1.0, 100000).astype(np.float32)
    return {
        "gamma_id_vs_ood": (g2, g3),                       # every bin resolves in the first pass
        "same_distribution": (g2, rng.gamma(2.0, 0.05, 150000).astype(np.float32)),
        "crossing_cdfs": (nrm, rng.normal(0.3, 2.0, 70000).astype(np.float32)),
        "zeros_vs_uniform": (np.zeros(1000, np.float32), rng.random(500).astype(np.float32)),
        "wide_range_signed_zero": (np.array([1e-10, 1e10, 3, -5], np.float32),
                                   np.array([2, -1e5, 0, -0.0, 7], np.float32)),
        "denormals": ((rng.random(4000) * 1e-39).astype(np.float32),
                      (rng.random(3000) * 2e-39 - 5e-40).astype(np.float32)),
        "single_values": (np.array([1.0], np.float32), np.array([2.0], np.float32)),
        "ties": (np.round(nrm, 1), np.round(nrm[::-1] * 1.5, 1).copy()),
        "tile_edges": (rng.random(2048 * 3).astype(np.float32),
                       (rng.random(2048 * 5 + 1) * 1.01).astype(np.float32)),
    }


@pytest.mark.parametrize("case", sorted(_binned_cases()))
def test_wasserstein_binned_method_equals_sort_method_and_oracle(case):
    """The one-pass bin-moment method, forced even where it has to sort most values, against the
    sort method (scipy's route), the float64 oracle, and the numpy stand-in of its bookkeeping."""
    u, v = _binned_cases()[case]
    ref = metrics_oracle.wasserstein_1d(u, v)
    cu, ku = metrics_oracle.bin_moments(u)
    cv, kv = metrics_oracle.bin_moments(v)
    book = metrics_oracle.bin_resolve(cu, ku, cv, kv)
    for a, b, swap in ((u, v, False), (v, u, True)):
        srt = ops.wasserstein_1d_info(_dev(a), _dev(b), "sort")
        bnd = ops.wasserstein_1d_info(_dev(a), _dev(b), "binned")
        auto = ops.wasserstein_1d_info(_dev(a), _dev(b), "auto")
        assert srt["method"] == "sort" and bnd["method"] == "binned"
        assert srt["value"] == pytest.approx(ref, rel=1e-11, abs=1e-300)
        assert bnd["value"] == pytest.approx(ref, rel=1e-11, abs=1e-300)
        assert auto["value"] == pytest.approx(ref, rel=1e-11, abs=1e-300)
        amb = (book["amb_v"], book["amb_u"]) if swap else (book["amb_u"], book["amb_v"])
        assert (bnd["sorted_u"], bnd["sorted_v"]) == amb
        expect_auto = "binned" if sum(amb) <= (a.size + b.size) // 2 else "sort"
        assert auto["method"] == expect_auto
    if case == "gamma_id_vs_ood":
        assert book["amb_u"] + book["amb_v"] == 0


def test_wasserstein_binned_unaligned_views_and_nonfinite_fallback():
    rng = np.random.default_rng(5)
    base_u = _dev(rng.gamma(2.0, 0.05, 100003).astype(np.float32))
    base_v = _dev(rng.gamma(3.0, 0.08, 100003).astype(np.float32))
    for off in (1, 2, 3):   # float4 body with a scalar head / tail
        u, v = base_u[off:], base_v[off + 1:-off]
        ref = metrics_oracle.wasserstein_1d(u.cpu().numpy(), v.cpu().numpy())
        assert ops.wasserstein_1d(u, v, "binned") == pytest.approx(ref, rel=1e-11)
    w = base_u.clone()
    w[17] = float("inf")
    info = ops.wasserstein_1d_info(w, base_v, "binned")
    assert info["method"] == "sort"            # inf / NaN: scipy's route decides what comes out
    assert info["value"] == ops.wasserstein_1d(w, base_v, "sort")
    # scipy: inf - inf inside the integral -> nan
    import math
    for bad_at in ((17,), (0, 5000, 100002)):
        wn = base_u.clone()
        wn[list(bad_at)] = float("nan")
        for m in ("auto", "sort"):
            assert math.isnan(ops.wasserstein_1d(wn, base_v, m)), (bad_at, m)   # as scipy: nan in, nan out
            assert math.isnan(ops.wasserstein_1d(base_v, wn, m)), (bad_at, m)
        assert math.isnan(ops.wasserstein_1d_async(wn, base_v).result())
    # and the library is still healthy afterwards
    ref = metrics_oracle.wasserstein_1d(base_u.cpu().numpy(), base_v.cpu().numpy())
    assert ops.wasserstein_1d(base_u, base_v, "sort") == pytest.approx(ref, rel=1e-11)


def test_wasserstein_binned_at_scale():
    """BASELINE configs[4] shape on one GPU (50 M + 50 M Gamma scores): one pass, nothing sorted,
    equal to the sort method to 1e-12."""
    n = 50_000_000
    g = torch.Generator(device=DEV).manual_seed(0)
    u = torch.empty(n, device=DEV).exponential_(1.0, generator=g)
    u = (u + torch.empty(n, device=DEV).exponential_(1.0, generator=g)) * 0.05       # Gamma(2, .05)
    v = torch.zeros(n, device=DEV)
    for _ in range(3):
        v += torch.empty(n, device=DEV).exponential_(1.0, generator=g)
    v *= 0.08                                                                         # Gamma(3, .08)
    b = ops.wasserstein_1d_info(u, v, "binned")
    s = ops.wasserstein_1d_info(u, v, "sort")
    assert b["method"] == "binned"
    assert b["sorted_u"] + b["sorted_v"] < n // 100
    assert b["value"] == pytest.approx(s["value"], rel=1e-12)
    assert b["value"] == pytest.approx(0.14, rel=0.01)   # E[v] - E[u] = 0.24 - 0.10 (v dominates u)


def test_wasserstein_large_shift_property():
    """size-independent property at scale: W1(u, u + c) == c, W1(u, u) == 0 (8 M + 8 M values)."""
    n = 8 * 1024 * 1024
    u = torch.rand(n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0)) * 0.25
    assert ops.wasserstein_1d(u, u) == 0.0
    assert ops.wasserstein_1d(u, u + 0.5) == pytest.approx(0.5, rel=2e-6)


@pytest.mark.parametrize("name", ["metrics_small.npz", "metrics_medium.npz"])
def test_kde_jsd_matches_reference_golden(name):
    g = load_golden(name)
    got = ops.kde_jsd(_dev(g["id"]), _dev(g["ood"]))
    ref = float(g["jsd"])
    print(f"[kde_jsd {name}] got {got:.12f} ref {ref:.12f} rel err {abs(got - ref) / ref:.3e}")
    assert got == pytest.approx(ref, rel=JSD_RTOL)


@pytest.mark.parametrize("nu,nv,grid", [(50, 60, 500), (2500, 700, 2000), (20000, 30000, 4096)])
def test_kde_jsd_against_oracle(nu, nv, grid):
    rng = np.random.default_rng(nu + nv)
    u = rng.gamma(2.0, 0.05, nu).astype(np.float32)
    v = rng.gamma(3.0, 0.08, nv).astype(np.float32)
    ref = metrics_oracle.pdf_jsd(u, v, num_points=grid)
    got = ops.kde_jsd(_dev(u), _dev(v), grid)
    print(f"[kde_jsd {nu}x{nv}x{grid}] rel err {abs(got - ref) / ref:.3e}")
    assert got == pytest.approx(ref, rel=JSD_RTOL)


def test_kde_jsd_properties():
    u = torch.rand(30000, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    assert ops.kde_jsd(u, u.clone()) == pytest.approx(0.0, abs=1e-6)
    v = u * 0.5 + 2.0  # disjoint supports -> distance -> sqrt(ln 2)
    d = ops.kde_jsd(u, v)
    assert d == pytest.approx(np.sqrt(np.log(2.0)), rel=1e-3)
    assert ops.kde_jsd(v, u) == pytest.approx(d, rel=1e-9)
    with pytest.raises(ValueError, match="at least 2 values"):
        ops.kde_jsd(u[:1], v)


def test_metric_classes_drop_in():
    """The evaluator mirror with a DummyModel as in reference tests/test_evaluation.py:11-28."""
    g = load_golden("metrics_small.npz")
    id_x, ood_x = torch.zeros(4, 3, device=DEV), torch.ones(4, 3, device=DEV)

    class DummyModel:
        def __call__(self, x, return_ue=False):
            s = g["id"] if float(x.sum()) == 0.0 else g["ood"]
            return None, torch.from_numpy(s).to(DEV).unsqueeze(-1)

        def eval(self):
            pass

    ev = evaluation.get_uncertainty_evaluator(["wasserstein_distance", "jensen_shannon_distance"])
    res = ev.evaluate(DummyModel(), (id_x, None), (ood_x, None))
    assert set(res) == {"wasserstein_distance", "jensen_shannon_distance"}
    assert all(isinstance(v, float) for v in res.values())
    assert res["wasserstein_distance"] == pytest.approx(float(g["wasserstein"]), rel=1e-12)
    assert res["jensen_shannon_distance"] == pytest.approx(float(g["jsd"]), rel=JSD_RTOL)
    two = evaluation.WassersteinEvaluation()._evaluate_uncertainties(
        evaluation.UncertaintyEstimate((torch.from_numpy(g["id"]), torch.from_numpy(g["id"]))),
        evaluation.UncertaintyEstimate((torch.from_numpy(g["ood"]), torch.from_numpy(g["ood"]))))
    assert two["wasserstein_distance"] == pytest.approx(float(g["wasserstein"]), rel=1e-12)


def _kde_cases():
    rng = np.random.default_rng(21)
    return {
        "gamma_id_vs_ood": (rng.gamma(2.0, 0.05, 9000), rng.gamma(3.0, 0.08, 7000)),
        "separated_normals": (rng.normal(0, 1, 8000), rng.normal(6, 0.5, 5000)),
        "bimodal_vs_wide": (np.concatenate([rng.normal(0, 0.1, 4000), rng.normal(5, 0.1, 4000)]),
                            rng.normal(2.5, 1, 6000)),
        "lognormal": (rng.lognormal(0, 1.0, 8000), rng.lognormal(0.5, 0.7, 8000)),
        "tiny": (rng.random(2), rng.random(3) + 0.5),
    }


@pytest.mark.parametrize("case", sorted(_kde_cases()))
def test_kde_jsd_moment_method_equals_window_method_and_oracle(case):
    """The one-pass moment method against the window method (every exp term within 9 bandwidths)
    and the float64 oracle of scipy's arithmetic; the two device methods agree far tighter than
    either does with float64 (their difference is the truncated Hermite series, <= 1e-8)."""
    u, v = (a.astype(np.float32) for a in _kde_cases()[case])
    G = 3000
    ref = metrics_oracle.pdf_jsd(u, v, G)
    mom = ops.kde_jsd_info(_dev(u), _dev(v), G, "moments")
    win = ops.kde_jsd_info(_dev(u), _dev(v), G, "window")
    auto = ops.kde_jsd_info(_dev(u), _dev(v), G, "auto")
    assert mom["method"] == "moments" and win["method"] == "window" and auto["method"] == "moments"
    print(f"[kde_jsd {case}] oracle {ref:.12f} moments {mom['value']:.12f} window {win['value']:.12f}")
    assert mom["value"] == pytest.approx(ref, rel=JSD_RTOL)
    assert win["value"] == pytest.approx(ref, rel=JSD_RTOL)
    assert mom["value"] == pytest.approx(win["value"], rel=2e-6)
    assert auto["value"] == pytest.approx(mom["value"], rel=1e-7)   # float atomics: order varies


def test_kde_jsd_moment_method_falls_back_when_bins_do_not_fit():
    rng = np.random.default_rng(2)
    u = rng.normal(0, 1, 200_000).astype(np.float32)
    u[0] = 3.0e4                      # one far outlier: (max - min) / (h / 4) >> 8192 bins
    v = rng.normal(0.5, 1, 150_000).astype(np.float32)
    info = ops.kde_jsd_info(_dev(u), _dev(v), 2000, "auto")
    assert info["method"] == "window" and np.isfinite(info["value"])
    with pytest.raises(ValueError, match="moment method needs"):
        ops.kde_jsd(_dev(u), _dev(v), 2000, "moments")
    with pytest.raises(ValueError, match="unknown KDE method"):
        ops.kde_jsd(_dev(u), _dev(v), 2000, "fft")


def test_kde_jsd_wide_range_uses_coarser_fine_bins():
    """(max - min) = 1400 bandwidths: the single-launch kernel (csrc/kde_fused.cu) has room for
    two fine bins per coarse bin instead of 64."""
    rng = np.random.default_rng(5)
    u = rng.normal(0, 1, 200_000).astype(np.float32)
    u[0] = 120.0
    v = rng.normal(0.5, 1, 150_000).astype(np.float32)
    mom = ops.kde_jsd_info(_dev(u), _dev(v), 2000, "moments")
    win = ops.kde_jsd_info(_dev(u), _dev(v), 2000, "window")
    assert mom["method"] == "moments" and win["method"] == "window"
    assert mom["value"] == pytest.approx(win["value"], rel=2e-6)


def test_kde_jsd_single_launch_is_reproducible_and_counts_one_launch():
    u, v = _gamma_pair(300_000, 200_000, seed=9)
    du, dv = _dev(u), _dev(v)
    first = ops.kde_jsd(du, dv, 20000)
    ops.reset_launch_count()
    again = ops.kde_jsd(du, dv, 20000)
    assert ops.launch_count() == 1          # one cooperative launch (plus a memset)
    assert again == first                   # fixed-order reductions, no float atomics
    ops.reset_launch_count()
    w = ops.wasserstein_1d(du, dv)
    assert ops.launch_count() == 1
    assert w == pytest.approx(metrics_oracle.wasserstein_1d(u, v), rel=1e-12)
    # unaligned views and odd sizes go through the same kernel
    ref = metrics_oracle.pdf_jsd(u[1:70_001], v[3:50_000], 1500)
    assert ops.kde_jsd(du[1:70_001], dv[3:50_000], 1500) == pytest.approx(ref, rel=JSD_RTOL)


def test_kde_jsd_methods_agree_at_scale():
    """BASELINE configs[4] shape on one GPU (50 M + 50 M, 20 000 grid points)."""
    n = 50_000_000
    g = torch.Generator(device=DEV).manual_seed(0)
    u = torch.empty(n, device=DEV).exponential_(1.0, generator=g)
    u = (u + torch.empty(n, device=DEV).exponential_(1.0, generator=g)) * 0.05
    v = torch.zeros(n, device=DEV)
    for _ in range(3):
        v += torch.empty(n, device=DEV).exponential_(1.0, generator=g)
    v *= 0.08
    mom = ops.kde_jsd_info(u, v, 20000, "moments")
    win = ops.kde_jsd_info(u, v, 20000, "window")
    assert mom["method"] == "moments" and win["method"] == "window"
    assert mom["value"] == pytest.approx(win["value"], rel=2e-6)


# ---- per-rank steps of the sharded metrics (each CUDA step against its numpy stand-in) ----------

def _gamma_pair(nu, nv, seed=3):
    rng = np.random.default_rng(seed)
    return (rng.gamma(2.0, 0.05, nu).astype(np.float32), rng.gamma(3.0, 0.08, nv).astype(np.float32))


def test_sample_stats_matches_numpy():
    u, _ = _gamma_pair(200_003, 10)
    mn, mx, mean, m2 = ops.sample_stats(torch.from_numpy(u).to(DEV))
    a = u.astype(np.float64)
    assert mn == a.min() and mx == a.max()
    assert abs(mean - a.mean()) <= 1e-12 * abs(a.mean())
    assert abs(m2 - ((a - a.mean()) ** 2).sum()) <= 1e-9 * ((a - a.mean()) ** 2).sum()


def test_key_histogram_and_partition_match_numpy():
    u, _ = _gamma_pair(300_001, 10)
    u[:1000] *= -1.0  # negative keys too
    t = torch.from_numpy(u).to(DEV)
    hist = ops.key_histogram(t).cpu().numpy()
    ref = np.bincount(metrics_oracle.key_bin(u), minlength=metrics_oracle.KEY_BINS)
    assert np.array_equal(hist, ref)
    from nnueehcs_b200.distributed import choose_bin_owners
    owners = choose_bin_owners(hist, 5)
    counts = [int(hist[owners == p].sum()) for p in range(5)]
    out = ops.partition_by_bin(t, torch.from_numpy(owners).to(DEV), counts).cpu().numpy()
    dst = owners[metrics_oracle.key_bin(u)]
    off = 0
    for p, c in enumerate(counts):
        seg = np.sort(out[off:off + c])
        assert np.array_equal(seg, np.sort(u[dst == p])), p
        off += c
    # ranges are ordered: every value of part p is <= every value of part p + 1
    bounds = np.cumsum(counts)
    for p in range(4):
        if counts[p] and counts[p + 1]:
            assert out[bounds[p] - counts[p]:bounds[p]].max() <= out[bounds[p]:bounds[p + 1]].min()


def test_wasserstein_ranges_compose_to_the_full_distance():
    u, v = _gamma_pair(120_000, 90_000)
    from nnueehcs_b200.distributed import choose_bin_owners
    hu = np.bincount(metrics_oracle.key_bin(u), minlength=metrics_oracle.KEY_BINS)
    hv = np.bincount(metrics_oracle.key_bin(v), minlength=metrics_oracle.KEY_BINS)
    owners = choose_bin_owners(hu + hv, 3)
    du, dv = owners[metrics_oracle.key_bin(u)], owners[metrics_oracle.key_bin(v)]
    total, prev, ub, vb = 0.0, None, 0, 0
    for p in range(3):
        up, vp = u[du == p], v[dv == p]
        part, first, last = ops.wasserstein_1d_range(torch.from_numpy(up).to(DEV),
                                                     torch.from_numpy(vp).to(DEV), ub, vb,
                                                     u.size, v.size)
        rp, rf, rl = metrics_oracle.wasserstein_1d_range(up, vp, ub, vb, u.size, v.size)
        assert abs(part - rp) <= 1e-10 * max(abs(rp), 1e-30) and first == rf and last == rl
        if prev is not None:
            total += abs(prev[1] - prev[2]) * (first - prev[0])
        total += part
        ub, vb = ub + up.size, vb + vp.size
        prev = (last, ub / u.size, vb / v.size)
    ref = metrics_oracle.wasserstein_1d(u, v)
    assert abs(total - ref) <= 1e-10 * ref


@pytest.mark.parametrize("case", ["gamma_id_vs_ood", "crossing_cdfs", "same_distribution",
                                  "wide_range_signed_zero"])
def test_binned_wasserstein_steps_match_numpy_stand_ins(case):
    """uq_bin_moments / uq_wasserstein_from_bins / uq_compact_flagged / uq_wasserstein_ambiguous
    (the per-rank steps of the sharded binned method) against oracle.metrics_oracle, shard-wise."""
    u, v = _binned_cases()[case]
    tables = torch.zeros((4, ops.key_bins()), dtype=torch.int64, device=DEV)
    cut_u, cut_v = u.size // 3, v.size // 2
    for part in (u[:cut_u], u[cut_u:]):          # two shards accumulate into the same tables
        ops.bin_moments(_dev(part), tables, 0)
    for part in (v[:cut_v], v[cut_v:]):
        ops.bin_moments(_dev(part), tables, 2)
    cu, ku = metrics_oracle.bin_moments(u)
    cv, kv = metrics_oracle.bin_moments(v)
    t = tables.cpu().numpy()
    assert np.array_equal(t[0], cu) and np.array_equal(t[1], ku)
    assert np.array_equal(t[2], cv) and np.array_equal(t[3], kv)
    book = metrics_oracle.bin_resolve(cu, ku, cv, kv)
    r = ops.wasserstein_from_bins(tables, u.size, v.size)
    assert (r["amb_u"], r["amb_v"], r["nonfinite"]) == (book["amb_u"], book["amb_v"], 0)
    assert np.array_equal(r["flags"].cpu().numpy(), book["flags"])
    assert r["resolved"] == pytest.approx(book["resolved"], rel=1e-12, abs=1e-300)
    au = torch.cat([ops.compact_flagged(_dev(u[:cut_u]), r["flags"]),
                    ops.compact_flagged(_dev(u[cut_u:]), r["flags"])])
    av = ops.compact_flagged(_dev(v), r["flags"])
    fl = book["flags"].astype(bool)
    assert np.array_equal(np.sort(au.cpu().numpy()), np.sort(u[fl[metrics_oracle.key_bin(u + np.float32(0))]]))
    assert av.numel() == book["amb_v"]
    ref = metrics_oracle.wasserstein_1d(u, v)
    if au.numel() + av.numel():
        exact = ops.wasserstein_ambiguous(au, av, tables, u.size, v.size)
        assert exact == pytest.approx(
            metrics_oracle.wasserstein_ambiguous(au.cpu().numpy(), av.cpu().numpy(), cu, ku, cv, kv),
            rel=1e-11, abs=1e-300)
        with pytest.raises(ValueError, match="the tables say"):
            ops.wasserstein_ambiguous(au[:-1], av, tables, u.size, v.size)
    else:
        exact = 0.0
    assert r["resolved"] + exact == pytest.approx(ref, rel=1e-11)


def test_kde_grid_accumulate_and_jsd_match_oracle():
    u, v = _gamma_pair(30_000, 20_000)
    G = 2000
    a, b = u.astype(np.float64), v.astype(np.float64)
    lo, hi = min(a.min(), b.min()), max(a.max(), b.max())
    hu = a.std(ddof=1) * a.size ** -0.2
    hv = b.std(ddof=1) * b.size ** -0.2
    grids = torch.zeros((2, G), dtype=torch.float64, device=DEV)
    # two shards per sample accumulate into the same grid
    ops.kde_grid_accumulate(torch.from_numpy(u[:11_000]).to(DEV), lo, hi, hu, grids[0])
    ops.kde_grid_accumulate(torch.from_numpy(u[11_000:]).to(DEV), lo, hi, hu, grids[0])
    ops.kde_grid_accumulate(torch.from_numpy(v).to(DEV), lo, hi, hv, grids[1])
    j = ops.jsd_from_grids(grids)
    ref = metrics_oracle.pdf_jsd(u, v, G)
    assert abs(j - ref) <= 2e-5 * ref, (j, ref)
    assert abs(j - ops.kde_jsd(torch.from_numpy(u).to(DEV), torch.from_numpy(v).to(DEV), G)) <= 1e-6 * ref


def test_sharded_metrics_single_rank_group():
    """world_size-1 NCCL group: the sharded entry points degenerate to the single-GPU kernels."""
    import torch.distributed as dist
    from nnueehcs_b200 import distributed as nd
    import os, socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=DEV)
    try:
        u, v = _gamma_pair(200_000, 150_000)
        tu, tv = torch.from_numpy(u).to(DEV), torch.from_numpy(v).to(DEV)
        w = nd.wasserstein_1d_sharded(tu, tv)
        j = nd.kde_jsd_sharded(tu, tv, 2000)
        w_ref = metrics_oracle.wasserstein_1d(u, v)
        j_ref = metrics_oracle.pdf_jsd(u, v, 2000)
        assert abs(w - w_ref) <= 1e-10 * w_ref
        assert abs(j - j_ref) <= 2e-5 * j_ref
        for m in ("binned", "sort"):
            info = {}
            assert abs(nd.wasserstein_1d_sharded(tu, tv, method=m, info=info) - w_ref) <= 1e-10 * w_ref
            assert info["method"] == m
        # crossing CDFs: some bins are ambiguous and their values go through the gather
        rng = np.random.default_rng(9)
        a = torch.from_numpy(rng.normal(0, 1, 90_000).astype(np.float32)).to(DEV)
        b = torch.from_numpy(rng.normal(0.3, 2, 60_000).astype(np.float32)).to(DEV)
        info = {}
        w2 = nd.wasserstein_1d_sharded(a, b, info=info)
        assert info["method"] == "binned" and 0 < info["exchanged_values"] < 40_000
        ref2 = metrics_oracle.wasserstein_1d(a.cpu().numpy(), b.cpu().numpy())
        assert abs(w2 - ref2) <= 1e-10 * ref2
    finally:
        if created:
            dist.destroy_process_group()


# ---- score consumers (mean / max / percentile score, AUROC, TNR@TPR, percentile classifier) ------

_SG = load_golden("score_metrics.npz")


@pytest.mark.parametrize("name", [str(n) for n in _SG["names"]])
def test_score_metrics_match_reference_golden(name):
    """uq_score_metrics against the outputs of the reference's own classes.  Counts-based metrics
    and AUROC are exact (integer arithmetic on the sorted arrays, one float64 division); the mean
    is a float64 sum here and a float32 pairwise sum in numpy -> 1e-6 relative."""
    id_s, ood_s = _SG[f"{name}.id"], _SG[f"{name}.ood"]
    a, b = _dev(id_s), _dev(ood_s)
    for i in range(len(_SG["tprs"])):
        q, t, p = float(_SG["pct_score"][i]), float(_SG["tprs"][i]), float(_SG["cls_pct"][i])
        for rev in (False, True):
            tag = "rev" if rev else "fwd"
            r = ops.score_metrics(a, b, percentile_q=q, target_tpr=t, tnr_reversed=rev,
                                  classifier_percentile=p, classifier_reversed=rev)
            assert r["tnr_at_tpr"] == float(_SG[f"{name}.tnr_{tag}"][i]), (name, t, rev)
            got = np.array([r["sensitivity"], r["specificity"], r["fpr"], r["fnr"]])
            assert np.array_equal(got, _SG[f"{name}.cls_{tag}"][i]), (name, p, rev, got)
            assert r["auroc"] == pytest.approx(float(_SG[f"{name}.auroc"]), rel=1e-14, abs=1e-15)
            assert r["max_score"] == float(_SG[f"{name}.max_score"])
            assert r["mean_score"] == pytest.approx(float(_SG[f"{name}.mean_score"]), rel=1e-6)
            assert r["percentile_score"] == pytest.approx(float(_SG[f"{name}.percentile_score"][i]),
                                                          rel=1e-12), (name, q)


def test_score_metrics_large_against_oracle_and_mirror_classes():
    u, v = _gamma_pair(400_000, 300_000, seed=9)
    a, b = _dev(u), _dev(v)
    r = ops.score_metrics(a, b, percentile_q=97.5, target_tpr=0.9, classifier_percentile=0.99)
    assert r["auroc"] == pytest.approx(metrics_oracle.auroc(u, v), rel=1e-14)
    assert r["tnr_at_tpr"] == metrics_oracle.tnr_at_tpr(u, v, 0.9)
    sens, spec, fpr, fnr = metrics_oracle.percentile_classifier(u, v, 0.99)
    assert (r["sensitivity"], r["specificity"], r["fpr"], r["fnr"]) == (sens, spec, fpr, fnr)
    mean, mx, pct = metrics_oracle.score_summaries(u, 97.5)
    assert r["max_score"] == mx and r["percentile_score"] == pytest.approx(pct, rel=1e-12)
    assert r["mean_score"] == pytest.approx(mean, rel=1e-6)
    # through the mirror classes (same names / keys as the reference)
    m = evaluation.TNRatTPX(0.9)
    assert m._evaluate_scores(a[:, None], b[:, None]) == {"tnr_at_tpr90": r["tnr_at_tpr"]}
    assert evaluation.AUROC()._evaluate_scores(a, b)["auroc"] == r["auroc"]
    c = evaluation.PercentileBasedClassifier(0.99)._evaluate_scores(a, b)
    assert c == {"sensitivity": sens, "specificity": spec}
    from nnueehcs_b200 import classification
    rv = classification.ReversedPercentileBasedIdOodClassifier(0.95)._evaluate_scores(a, b)
    thr = float(np.quantile(u.astype(np.float64), 0.05))
    assert abs(rv["sensitivity"] - float((v <= thr).mean())) < 1e-4
    assert abs(rv["specificity"] - float((u > thr).mean())) < 1e-4
    with pytest.raises(ValueError, match="between 0 and 1"):
        ops.score_metrics(a, b, target_tpr=2.0)


# ---- KDEMLPModel's input-density score --------------------------------------------

KDE_RTOL = 2e-5   # sklearn's own tree pruning is rtol 1e-5; the kernel sums float32 MUFU.EX2 terms


@pytest.mark.parametrize("tag", ["binomial5", "d2_half", "d12"])
def test_kde_density_matches_reference_class_golden(tag):
    g = load_golden("kde_density.npz")
    fit, x, h = g[f"{tag}.fit"], g[f"{tag}.x"], float(g[f"{tag}.bandwidth"])
    assert ops.kde_scott_bandwidth(*fit.shape) == pytest.approx(h, rel=1e-15)
    got = ops.kde_density(_dev(fit), _dev(x), h)
    assert got.dtype == torch.float64 and tuple(got.shape) == (x.shape[0],)
    got = got.cpu().numpy()
    ref = metrics_oracle.kde_neg_density(fit, x, h)
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    print(f"[kde_density {tag}] max rel err vs float64 oracle {np.max(err / np.maximum(np.abs(ref), 1e-300)):.2e}")
    assert np.all(err <= KDE_RTOL * np.abs(ref) + 1e-12 * scale)
    assert np.all(np.abs(got - g[f"{tag}.dens"]) <= 2 * KDE_RTOL * np.abs(ref) + 1e-12 * scale)
    assert got[-1] == 0.0          # far query: every term underflows, as in the reference


@pytest.mark.parametrize("tag", ["far5", "far2", "silverman3"])
def test_kde_density_far_queries_keep_their_order(tag):
    """Far-OOD queries (reference densities 1e-38 ... 1e-318, where a float32 kernel sum has
    flushed to zero) are re-done in float64 log space: equal to the float64 oracle at the same
    tolerance as near queries, to the reference's own KDEMLPModel within ITS far-field error
    (tree pruning, 6e-5 measured), and ranked like the reference ranks them."""
    g = load_golden("kde_density_far.npz")
    fit, x, h = g[f"{tag}.fit"], g[f"{tag}.x"], float(g[f"{tag}.bandwidth"])
    got = ops.kde_density(_dev(fit), _dev(x), h).cpu().numpy()
    ref = metrics_oracle.kde_neg_density(fit, x, h)
    gold = g[f"{tag}.dens"]
    np.testing.assert_allclose(got, ref, rtol=KDE_RTOL, atol=1e-300)
    np.testing.assert_allclose(got, gold, rtol=2e-4, atol=1e-300)
    assert np.array_equal(got == 0, gold == 0)
    big = np.abs(gold) > 1e-290                  # away from float64 denormals: strict order kept
    assert np.array_equal(np.argsort(got[big], kind="stable"), np.argsort(gold[big], kind="stable"))


def test_kde_wrapper_silverman_bandwidth():
    """bandwidth='silverman' (the reference's KDE search space offers both rules,
    examples/bo_driven/config_kde.yaml:385-390) through the wrapper, against the reference class."""
    from nnueehcs_b200.model_builder import KDEModelBuilder
    g = load_golden("kde_density_far.npz")
    fit, x = g["silverman3.fit"], g["silverman3.x"]
    arch = [{"Linear": {"args": [3, 16]}}, {"ReLU": {"inplace": True}}, {"Linear": {"args": [16, 1]}}]
    model = KDEModelBuilder(arch, {"bandwidth": "silverman", "rtol": 0.1, "train_fit_prop": 1.0}).build()
    model.to(DEV).eval()
    model.fit_kde(torch.from_numpy(fit).to(DEV))
    assert model.kde["bandwidth_"] == pytest.approx(float(g["silverman3.bandwidth"]), rel=1e-15)
    assert ops.kde_silverman_bandwidth(*fit.shape) == pytest.approx(
        metrics_oracle.silverman_bandwidth_sklearn(*fit.shape), rel=1e-15)
    with torch.no_grad():
        _, dens = model(torch.from_numpy(x).to(DEV), return_ue=True)
    np.testing.assert_allclose(dens.cpu().numpy(), g["silverman3.dens"], rtol=2e-4, atol=1e-300)
    with pytest.raises(ValueError, match="'scott', 'silverman' or a number"):
        KDEModelBuilder(arch, {"bandwidth": "epanechnikov"}).build()


def test_kde_density_fit_splits_ragged_sizes_and_errors():
    """Few query rows x many fitted rows (the fitted rows are split over blocks), ragged tiles."""
    rng = np.random.default_rng(4)
    for n, m, d in ((7, 70_001, 5), (1025, 4097, 3), (1, 1, 8), (3000, 513, 1)):
        fit = rng.random((m, d)).astype(np.float32)
        x = (rng.random((n, d)) * 1.5 - 0.25).astype(np.float32)
        h = metrics_oracle.scott_bandwidth_sklearn(m, d)
        got = ops.kde_density(_dev(fit), _dev(x), h).cpu().numpy()
        ref = metrics_oracle.kde_neg_density(fit, x, h)
        assert np.all(np.abs(got - ref) <= KDE_RTOL * np.abs(ref) + 1e-12 * np.abs(ref).max()), (n, m, d)
    a = torch.rand(10, 5, device=DEV)
    with pytest.raises(ValueError, match="agree in d"):
        ops.kde_density(a, a[:, :4], 0.3)
    with pytest.raises(ValueError, match="at least one"):
        ops.kde_density(a[:0], a, 0.3)
    with pytest.raises(ValueError, match="supported: 1..32"):
        ops.kde_density(torch.rand(10, 40, device=DEV), torch.rand(4, 40, device=DEV), 0.3)
    with pytest.raises(ValueError, match="bandwidth must be positive"):
        ops.kde_density(a, a, 0.0)


def test_kde_wrapper_drop_in():
    from nnueehcs_b200.model_builder import KDEModelBuilder
    import yaml as _yaml
    g = load_golden("kde_density.npz")
    tag = "binomial5"
    arch = _yaml.safe_load(str(g[f"{tag}.arch_yaml"]))
    model = KDEModelBuilder(arch, {"bandwidth": "scott", "rtol": 0.1, "train_fit_prop": 1.0}).build()
    model.model.load_state_dict({k[len(f"{tag}.m0."):]: torch.from_numpy(g[k]) for k in g.files
                                 if k.startswith(f"{tag}.m0.")})
    model.to(DEV)
    model.eval()
    fit = torch.from_numpy(g[f"{tag}.fit"]).to(DEV)
    model.fit_kde(fit)                       # train_fit_prop = 1: a permutation of the same rows
    assert model.kde["bandwidth_"] == pytest.approx(float(g[f"{tag}.bandwidth"]), rel=1e-15)
    with torch.no_grad():
        pred, dens = model(torch.from_numpy(g[f"{tag}.x"]).to(DEV), return_ue=True)
    np.testing.assert_allclose(pred.cpu().numpy(), g[f"{tag}.pred"], rtol=1e-5, atol=1e-6)
    ref = g[f"{tag}.dens"]
    assert np.all(np.abs(dens.cpu().numpy() - ref) <= 2 * KDE_RTOL * np.abs(ref) + 1e-12)
    # a state_dict round trip keeps the fitted rows (they are a buffer)
    assert "_kde_data" in model.state_dict()


def test_kde_density_property_at_scale():
    """1 M queries x 100 k fitted rows (1e11 Gaussian terms): the density of the fitted sample
    integrates to ~1 over its support and a permutation of the fitted rows changes nothing
    beyond float64 summation order."""
    g = torch.Generator(device=DEV).manual_seed(0)
    fit = torch.rand(100_000, 5, device=DEV, generator=g)
    x = torch.rand(1 << 20, 5, device=DEV, generator=g) * 3.0 - 1.0    # uniform on [-1, 2]^5
    h = ops.kde_scott_bandwidth(*fit.shape)
    d1 = ops.kde_density(fit, x, h)
    d2 = ops.kde_density(fit[torch.randperm(fit.shape[0], device=DEV)], x, h)
    assert float((d1 - d2).abs().max()) <= 1e-6 * float(d1.abs().max())
    integral = float(-d1.mean()) * 3.0 ** 5          # Monte-Carlo integral over the box
    assert integral == pytest.approx(1.0, abs=0.05)   # Monte-Carlo error of 1 M samples ~ 1 %


def test_wasserstein_methods_on_random_bit_patterns():
    """Every finite float32 bit pattern is fair game: both signs, every exponent, denormals.  The
    binned method's edge / ulp decoding is exercised on all 16 320 finite key bins."""
    rng = np.random.default_rng(12)

    def rand_bits(n):
        f = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
        return f[np.isfinite(f)]

    for trial in range(3):
        u, v = rand_bits(200_000 + 17 * trial), rand_bits(150_000)
        if trial == 2:   # clusters that differ only far below / above the bulk
            u = np.concatenate([u, rng.normal(0, 1e-30, 50_000).astype(np.float32),
                                rng.normal(5, 1e-3, 50_000).astype(np.float32)])
            v = np.concatenate([v, rng.normal(0, 2e-30, 70_000).astype(np.float32),
                                rng.normal(5, 2e-3, 30_000).astype(np.float32)])
        ref = metrics_oracle.wasserstein_1d(u, v)
        for method in ("sort", "binned", "auto"):
            got = ops.wasserstein_1d(_dev(u), _dev(v), method)
            assert got == pytest.approx(ref, rel=1e-11), (trial, method)


# ---- the radix sort on its own ----------------------------------------------------------------------

def _sort_ref(x):
    """np.sort on the order-preserving integer keys (bit-pattern order: -0.0 < +0.0, NaNs last)."""
    b = x.view(np.uint32)
    k = np.where(b >> 31, ~b, b | np.uint32(0x80000000)).astype(np.uint32)
    k.sort()
    return np.where(k >> 31, k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32)


@pytest.mark.parametrize("n", [1, 2, 31, 33, 4095, 4096, 4097, 8191, 8193, 70001, 1 << 20,
                               2_424_833, 2_424_832 + 8192 * 3 + 17])
def test_radix_sort_bit_exact(n):
    rng = np.random.default_rng(n)
    kinds = [rng.standard_normal(n).astype(np.float32),
             rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32),
             np.round(rng.gamma(2.0, 0.05, n), 2).astype(np.float32),        # heavy ties
             np.full(n, 1.5, np.float32)]                                     # one digit everywhere
    kinds[1][np.isnan(kinds[1])] = np.float32(np.inf)   # NaN payloads: compare everything else
    for x in kinds:
        if n > 3:
            x[:3] = np.array([0.0, -0.0, -np.inf], np.float32)
        got = ops.sort_f32(_dev(x)).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), _sort_ref(x))
    if n > 8:   # a slice that is not 16-byte aligned (copied before pass 0 instead of read in place)
        x = kinds[0]
        got = ops.sort_f32(_dev(x)[1:]).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), _sort_ref(x[1:].copy()))


def test_radix_sort_puts_every_nan_last():
    """np.sort's convention: NaNs of either sign and any payload after +inf (they come back as one
    canonical NaN); everything else in bit-pattern order."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal(100_003).astype(np.float32)
    bits = x.view(np.uint32)
    bits[5] = 0x7FC00000          # quiet NaN
    bits[77] = 0xFFC00000         # negative quiet NaN
    bits[4096] = 0x7F800001       # signalling NaN, small payload
    bits[99_999] = 0xFFFFFFFF     # negative NaN, full payload
    x[10], x[11] = np.inf, -np.inf
    got = ops.sort_f32(_dev(x)).cpu().numpy()
    assert np.isnan(got[-4:]).all() and not np.isnan(got[:-4]).any()
    assert np.array_equal(got[:-4].view(np.uint32), _sort_ref(x[~np.isnan(x)].copy()))
    assert got[-5] == np.inf and got[0] == -np.inf


def test_radix_sort_at_baseline_size_properties():
    """50 M keys (BASELINE configs[4]'s per-sample size), beyond what numpy sorts in seconds:
    size-independent properties -- sorted, the same multiset (sum and xor of the bit patterns), idempotent --
    and equality with torch.sort of the same GPU (a library sort, the checker here)."""
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(50_000_000, generator=g, device=DEV) * torch.rand(50_000_000, generator=g, device=DEV)
    x[:5] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 1e-42], device=DEV)
    y = ops.sort_f32(x)
    assert bool((y[1:] >= y[:-1]).all())
    xi, yi = x.view(torch.int32), y.view(torch.int32)
    assert int(xi.to(torch.int64).sum()) == int(yi.to(torch.int64).sum())
    def xor_all(t):
        while t.numel() > 1:
            h = t.numel() // 2
            t = torch.cat([t[:h] ^ t[h:2 * h], t[2 * h:]])
        return int(t[0])
    assert xor_all(xi.clone()) == xor_all(yi.clone())
    assert torch.equal(ops.sort_f32(y).view(torch.int32), yi)
    ref = torch.sort(x).values
    assert torch.equal(ref, y)          # values equal everywhere (-0.0 == +0.0 under this comparison)


def test_radix_sort_both_tile_sizes():
    """8192-key tiles (32 keys per lane) and 4096-key tiles give the same bits; UQ_SORT_ITEMS is
    read once per process, so the forced sizes run in subprocesses."""
    import subprocess
    import sys
    code = ("import numpy as np, torch, zlib; from nnueehcs_b200 import ops;"
            "x = torch.from_numpy(np.random.default_rng(5).standard_normal(3_000_017)"
            ".astype(np.float32)).cuda();"
            "print(zlib.crc32(ops.sort_f32(x).cpu().numpy().tobytes()))")
    import os
    outs = []
    for items in ("16", "32"):
        env = dict(os.environ, UQ_SORT_ITEMS=items)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip().splitlines()[-1])
    x = np.random.default_rng(5).standard_normal(3_000_017).astype(np.float32)
    import zlib
    assert outs[0] == outs[1] == str(zlib.crc32(_sort_ref(x).tobytes()))


# ---- enqueue / finish forms ---------------------------------------------------------------------------

def test_enqueued_metrics_equal_the_synchronous_calls():
    """uq_*_enqueue / uq_*_finish: several metrics in flight behind one synchronisation, the same
    kernels -> the same bits; inputs the single-launch kernels do not cover fall back inside
    finish (same-distribution samples: every bin ambiguous; inf: scipy's route)."""
    rng = np.random.default_rng(21)
    u = _dev(rng.gamma(2.0, 0.05, 300_001).astype(np.float32))
    v = _dev(rng.gamma(3.0, 0.08, 200_003).astype(np.float32))
    same = _dev(rng.gamma(2.0, 0.05, 250_000).astype(np.float32))
    pend = [ops.wasserstein_1d_async(u, v), ops.kde_jsd_async(u, v, 20000),
            ops.wasserstein_1d_async(u, same), ops.kde_jsd_async(v, same, 5000),
            ops.wasserstein_1d_async(v, u)]
    got = [p.result() for p in pend]
    assert got[0] == ops.wasserstein_1d(u, v)
    assert got[1] == ops.kde_jsd(u, v, 20000)
    assert got[2] == ops.wasserstein_1d(u, same)          # fallback: ambiguous bins
    assert got[3] == ops.kde_jsd(v, same, 5000)
    assert got[4] == got[0]
    assert got[0] == pytest.approx(metrics_oracle.wasserstein_1d(u.cpu().numpy(), v.cpu().numpy()),
                                   rel=1e-11)
    assert pend[0].result() == got[0]                      # idempotent
    w = u.clone()
    w[5] = float("inf")
    assert ops.wasserstein_1d_async(w, v).result() == ops.wasserstein_1d(w, v)
    wide = torch.cat([u, u.new_tensor([1e9])])            # range beyond 4000 bandwidths: KDE declines
    assert ops.kde_jsd_async(wide, v, 2000).result() == ops.kde_jsd(wide, v, 2000)
    with pytest.raises(ValueError, match="can't be empty"):
        ops.wasserstein_1d_async(u[:0], v)
    # the record must be mapped pinned host memory: pageable memory is refused, not written to
    import ctypes
    from nnueehcs_b200 import _lib
    lib = _lib.load()
    wsb = int(lib.uq_wasserstein_workspace_bytes(u.numel(), v.numel()))
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    pageable = (ctypes.c_uint8 * 256)()
    rc = lib.uq_wasserstein_1d_enqueue(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(),
                                       ctypes.addressof(pageable), ws.data_ptr(), wsb,
                                       torch.cuda.current_stream().cuda_stream)
    assert rc == _lib.UQ_ERR_INVALID and b"pinned" in lib.uq_last_error()
    torch.cuda.synchronize()
    # finish before the kernel has written the record is an error, not a wrong number
    rec = torch.zeros(256, dtype=torch.uint8, pin_memory=True)
    rec.view(torch.int64)[3] = -1        # what enqueue writes before the launch
    out = ctypes.c_double()
    rc = lib.uq_wasserstein_1d_finish(u.data_ptr(), u.numel(), v.data_ptr(), v.numel(),
                                      rec.data_ptr(), ctypes.byref(out), None, ws.data_ptr(), wsb,
                                      torch.cuda.current_stream().cuda_stream)
    assert rc == _lib.UQ_ERR_INVALID and b"synchronise" in lib.uq_last_error()


def test_metric_evaluator_reads_distance_metrics_behind_one_synchronisation():
    class Scores(torch.nn.Module):
        def forward(self, x, return_ue=False):
            return x, x[:, 0].abs()

    rng = np.random.default_rng(3)
    xi = _dev(rng.gamma(2.0, 0.05, (50_000, 1)).astype(np.float32))
    xo = _dev(rng.gamma(3.0, 0.08, (40_000, 1)).astype(np.float32))
    ev = evaluation.get_uncertainty_evaluator(["wasserstein_distance", "jensen_shannon_distance",
                                               "auroc"])
    r = ev.evaluate(Scores(), (xi, None), (xo, None))
    assert r["wasserstein_distance"] == ops.wasserstein_1d(xi[:, 0], xo[:, 0])
    assert r["jensen_shannon_distance"] == ops.kde_jsd(xi[:, 0], xo[:, 0], 20000)
    assert isinstance(r["wasserstein_distance"], float) and "auroc" in r
    assert list(r)[:2] == ["wasserstein_distance", "jensen_shannon_distance"]


def test_runtime_throughput_and_memory_metrics_on_a_uq_model():
    """The reference's own throughput definition (evaluation.py:494-516) and its memory metric
    (:383-411) over the fused MC-dropout forward."""
    from nnueehcs_b200 import model_builder as mb
    arch = [{"Linear": {"args": [5, 128]}}, {"ReLU": {"inplace": True}},
            {"Linear": {"args": [128, 128]}}, {"ReLU": {"inplace": True}}, {"Linear": {"args": [128, 1]}}]
    torch.manual_seed(0)
    model = mb.MCDropoutModelBuilder(arch, {"num_samples": 20, "dropout_percent": 0.2}).build().to(DEV)
    idd, ood = (torch.rand(3000, 5, device=DEV), None), (torch.rand(2000, 5, device=DEV) + 0.5, None)
    ev = evaluation.get_evaluator([{"name": "uncertainty_estimating_throughput", "trials": 3, "warmup": 2},
                                   {"name": "uncertainty_estimating_runtime", "trials": 2, "warmup": 0},
                                   {"name": "max_memory_usage"}, {"name": "wasserstein"}])
    r = ev.evaluate(model, idd, ood)
    assert r["uncertainty_estimating_throughput"] > 1e4 and r["throughput_std"] >= 0
    assert 0 < r["runtime"] < 1.0 and r["max_memory_usage"] > 0
    assert r["wasserstein_distance"] >= 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process():
    """dynamic shared-memory opt-ins are per device.  Every kernel family runs on
    cuda:0 first and then on cuda:1 in the same process; the results must be identical."""
    from nnueehcs_b200 import model_builder as mb
    rng = np.random.default_rng(11)
    u = rng.gamma(2.0, 0.05, 3_000_017).astype(np.float32)
    v = rng.gamma(3.0, 0.08, 2_500_003).astype(np.float32)
    arch = [{"Linear": {"args": [5, 128]}}, {"BatchNorm1d": {"args": [128]}}, {"ReLU": {"inplace": True}},
            {"Linear": {"args": [128, 128]}}, {"ReLU": {"inplace": True}}, {"Linear": {"args": [128, 1]}}]
    torch.manual_seed(3)
    model = mb.EnsembleModelBuilder(arch, {"num_models": 4}).build().eval()
    x = torch.rand(5000, 5)
    out = []
    for idx in (0, 1):
        dev = torch.device("cuda", idx)
        ud, vd = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
        with torch.cuda.device(dev):
            pend = [ops.wasserstein_1d_async(ud, vd), ops.kde_jsd_async(ud, vd, 5000)]
            res = [ops.wasserstein_1d(ud, vd), ops.wasserstein_1d(ud, vd, "sort"),
                   ops.kde_jsd(ud, vd, 5000), ops.kde_jsd(ud[:200_000], vd[:200_000], 2000, "window"),
                   ops.score_metrics(ud, vd)["auroc"], [p.result() for p in pend],
                   ops.sort_f32(ud).cpu().numpy().tobytes(),
                   ops.kde_density(torch.rand(3000, 5, generator=torch.Generator().manual_seed(1)).to(dev),
                                   x[:512].to(dev), 0.3).cpu().numpy().tobytes()]
            model.to(dev)
            with torch.no_grad():
                for prec in ("fp32", "bf16"):
                    model.uq_precision = prec
                    mean, std = model(x.to(dev), return_ue=True)
                    res.append((mean.cpu().numpy().tobytes(), std.cpu().numpy().tobytes()))
        out.append(res)
    # the window KDE adds into its grid with float64 atomics: last-digit differences between runs
    assert out[0][3] == pytest.approx(out[1][3], rel=1e-12)
    out[0][3] = out[1][3] = None
    assert out[0] == out[1]
