#!/usr/bin/env python
"""Golden fixtures for the score consumers: outputs of the
reference's OWN metric classes -- ``MeanScoreEvaluation`` / ``MaxScoreEvaluation`` /
``PercentileScoreEvaluation`` (nnueehcs/evaluation.py:292-381), ``TNRatTPX`` (:519-605),
``AUROC`` (:607-635, sklearn ``roc_auc_score``) and ``PercentileBasedClassifier`` (:637-662 over
``classification.py:103-143``, ``torch.quantile``) -- on seeded score vectors with and without
ties.  Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden_scores.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import shims  # noqa: E402

_, _, ref_eval = shims.import_reference()

PCT_SCORE = [95.0, 50.0, 99.9, 0.0, 100.0]
TPRS = [0.95, 0.5, 0.99, 0.0, 1.0]
CLS_PCT = [0.95, 0.5, 0.999, 0.0, 1.0]


def cases():
    rng = np.random.default_rng(11)
    yield "gamma", rng.gamma(2.0, 0.05, 4000).astype(np.float32), rng.gamma(3.0, 0.08, 3000).astype(np.float32)
    yield "ties", (rng.integers(0, 40, 2500) * 0.25).astype(np.float32), (rng.integers(10, 60, 1800) * 0.25).astype(np.float32)
    yield "separated", rng.uniform(0, 1, 500).astype(np.float32), rng.uniform(2, 3, 700).astype(np.float32)
    yield "inverted", rng.uniform(2, 3, 600).astype(np.float32), rng.uniform(0, 1, 400).astype(np.float32)
    yield "constant_id", np.full(300, 0.5, np.float32), rng.uniform(0, 1, 200).astype(np.float32)
    yield "tiny", np.array([0.3, 0.1, 0.2], np.float32), np.array([0.25, 0.4], np.float32)


def main():
    out = {"pct_score": np.array(PCT_SCORE), "tprs": np.array(TPRS), "cls_pct": np.array(CLS_PCT)}
    names = []
    for name, id_s, ood_s in cases():
        names.append(name)
        tid, tood = torch.from_numpy(id_s)[:, None], torch.from_numpy(ood_s)[:, None]
        ue_id, ue_ood = ref_eval.UncertaintyEstimate(tid), ref_eval.UncertaintyEstimate(tood)
        out[f"{name}.id"], out[f"{name}.ood"] = id_s, ood_s
        out[f"{name}.mean_score"] = np.float64(
            ref_eval.MeanScoreEvaluation()._evaluate_uncertainties(ue_id, ue_ood)["mean_score"])
        out[f"{name}.max_score"] = np.float64(
            ref_eval.MaxScoreEvaluation()._evaluate_uncertainties(ue_id, ue_ood)["max_score"])
        out[f"{name}.percentile_score"] = np.array(
            [float(ref_eval.PercentileScoreEvaluation(q)._evaluate_uncertainties(ue_id, ue_ood)["percentile_score"])
             for q in PCT_SCORE])
        out[f"{name}.auroc"] = np.float64(ref_eval.AUROC()._evaluate_scores(tid, tood)["auroc"])
        for rev in (False, True):
            tag = "rev" if rev else "fwd"
            vals = []
            for t in TPRS:
                m = ref_eval.TNRatTPX(t, rev)
                vals.append(float(m._evaluate_scores(tid, tood)[str(m)]))
            out[f"{name}.tnr_{tag}"] = np.array(vals)
            cls = []
            for p in CLS_PCT:
                r = ref_eval.PercentileBasedClassifier(p, rev)._evaluate_scores(tid, tood)
                full = ref_eval.PercentileBasedClassifier(p, rev)._classifier._evaluate_scores(
                    -tid if rev else tid, -tood if rev else tood)
                assert r["sensitivity"] == full["sensitivity"]
                cls.append([full["sensitivity"], full["specificity"], full["fpr"], full["fnr"]])
            out[f"{name}.cls_{tag}"] = np.array(cls)
        print(name, "auroc", out[f"{name}.auroc"], "tnr", out[f"{name}.tnr_fwd"], "cls95",
              out[f"{name}.cls_fwd"][0], "pct", out[f"{name}.percentile_score"][:3])
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "score_metrics.npz"), **out)


if __name__ == "__main__":
    main()
