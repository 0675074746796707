"""The score-consumer oracle (closed forms on sorted arrays) against the goldens produced by the
reference's own metric classes (tests/golden/make_golden_scores.py)."""
import numpy as np
import pytest

from oracle import metrics_oracle as mo
from tests.util import load_golden

G = load_golden("score_metrics.npz")
NAMES = [str(n) for n in G["names"]]


@pytest.mark.parametrize("name", NAMES)
def test_score_oracle_matches_reference_golden(name):
    id_s, ood_s = G[f"{name}.id"], G[f"{name}.ood"]
    assert mo.auroc(id_s, ood_s) == pytest.approx(float(G[f"{name}.auroc"]), rel=1e-12, abs=1e-15)
    for i, t in enumerate(G["tprs"]):
        assert mo.tnr_at_tpr(id_s, ood_s, float(t), False) == pytest.approx(float(G[f"{name}.tnr_fwd"][i]), abs=0), (name, t)
        assert mo.tnr_at_tpr(id_s, ood_s, float(t), True) == pytest.approx(float(G[f"{name}.tnr_rev"][i]), abs=0), (name, t, "rev")
    for i, p in enumerate(G["cls_pct"]):
        assert np.array_equal(np.array(mo.percentile_classifier(id_s, ood_s, float(p), False)), G[f"{name}.cls_fwd"][i]), (name, p)
        assert np.array_equal(np.array(mo.percentile_classifier(id_s, ood_s, float(p), True)), G[f"{name}.cls_rev"][i]), (name, p, "rev")
    for i, q in enumerate(G["pct_score"]):
        mean, mx, pct = mo.score_summaries(id_s, float(q))
        assert mean == float(G[f"{name}.mean_score"]) and mx == float(G[f"{name}.max_score"])
        assert pct == float(G[f"{name}.percentile_score"][i])


def test_tnr_closed_form_equals_the_reference_loop_on_random_small_cases():
    rng = np.random.default_rng(3)
    for _ in range(200):
        n_id, n_ood = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        a = rng.integers(0, 6, n_id).astype(np.float32)
        b = rng.integers(2, 9, n_ood).astype(np.float32)
        for target in (0.0, 0.3, 0.5, 0.95, 1.0):
            for rev in (False, True):
                # literal restatement of nnueehcs/evaluation.py:538-580
                if (rev and a.min() > b.max()) or (not rev and a.max() < b.min()):
                    ref = 1.0
                else:
                    ref = 0.0
                    for thr in np.unique(np.concatenate([a, b])):
                        tp = int((a > thr).sum()) if rev else int((b > thr).sum())
                        tn = int((b <= thr).sum()) if rev else int((a <= thr).sum())
                        tpr, tnr = tp / n_ood, tn / n_id
                        if tpr >= target and tnr > ref:
                            ref = tnr
                assert mo.tnr_at_tpr(a, b, target, rev) == ref, (a, b, target, rev)
