#!/usr/bin/env python
"""Summarise a UQ_TC_TRACE timeline (CSV: role, kind, index, clock) of CTA 0."""
import collections
import csv
import sys


def main(path, first=2, last=5):
    rows = [tuple(map(int, r)) for r in csv.reader(open(path))]
    t0 = min(r[3] for r in rows)
    ev = collections.defaultdict(dict)
    for role, kind, idx, clk in rows:
        ev[(role, kind)][idx] = clk - t0
    dfull = ev[(2, 1)]
    mma_layer = ev[(1, 4)]
    print("layer-step g: MMA layer commit issued | epi(w2) d_full seen | epi w2 chunk-done times | epi w6 chunk-done")
    gs = sorted(dfull)
    for g in gs[first * 3:last * 3]:
        c2 = [ev[(2, 2)].get((g << 4) | c) for c in range(0, 8, 2)]
        c6 = [ev[(2, 6)].get((g << 4) | c) for c in range(1, 8, 2)]
        print(g, mma_layer.get(g), dfull.get(g), ev[(2, 5)].get(g), c2, c6)
    # per-stage MMA timeline statistics
    mb, mo, md = ev[(1, 1)], ev[(1, 2)], ev[(1, 3)]
    its = sorted(md)
    gaps = [md[i] - md[i - 1] for i in its[1:] if i - 1 in md]
    waits = [mo[i] - mb[i] for i in its if i in mo and i in mb]
    issue = [md[i] - mo[i] for i in its if i in mo]
    import statistics as st
    print("stages traced", len(its))
    print("stage period  median %.0f  p10 %.0f  p90 %.0f" % (st.median(gaps), sorted(gaps)[len(gaps) // 10], sorted(gaps)[9 * len(gaps) // 10]))
    print("w_full wait   median %.0f  mean %.0f" % (st.median(waits), st.mean(waits)))
    print("issue 4 MMAs+commit median %.0f" % st.median(issue))
    pi = ev[(0, 2)]
    pits = sorted(pi)
    pg = [pi[i] - pi[i - 1] for i in pits[1:]]
    print("producer issue period median %.0f" % st.median(pg))
    # item time
    if len(gs) > 9:
        print("cycles per member (3 layer-steps): %.0f" % ((dfull[gs[-1]] - dfull[gs[0]]) / (len(gs) - 1) * 3))


if __name__ == "__main__":
    main(sys.argv[1])
