#!/usr/bin/env python
"""Timeline summary of a UQ_TC_TRACE capture of the CTA-pair kernel (mlp_tc2.cu).
Roles: 0 producer(leader) 1 MMA 2 leader epi warp2 3 leader epi warp6 4 peer epi warp2 5 peer relay."""
import collections, csv, statistics as st, sys

def main(path, g0=12, g1=21):
    rows = [tuple(map(int, r)) for r in csv.reader(open(path))]
    t0 = min(r[3] for r in rows)
    ev = collections.defaultdict(dict)
    for role, kind, idx, clk in rows:
        ev[(role, kind)][idx] = clk - t0
    mb, mo, md = ev[(1, 1)], ev[(1, 2)], ev[(1, 3)]
    its = sorted(md)
    gaps = [md[i] - md[i - 1] for i in its[1:] if i - 1 in md]
    print("stage period median %.0f mean %.0f" % (st.median(gaps), st.mean(gaps)))
    commit, half0 = ev[(1, 4)], ev[(1, 5)]
    print("g | commit | +half0 chunks ready (MMA resumes, rel. to prev commit) | leader w2: top dfull done | leader w6 | peer w2   (relative to commit[g])")
    for g in range(g0, g1):
        c = commit.get(g)
        if c is None: continue
        def rel(role, kind):
            v = ev[(role, kind)].get(g)
            return None if v is None else v - c
        nxt = half0.get(g + 1)
        print(g, c, (nxt - c) if nxt else None, "|", rel(2, 0), rel(2, 1), rel(2, 2), "|", rel(3, 0), rel(3, 1), rel(3, 2),
              "|", rel(4, 0), rel(4, 1), rel(4, 2), "| next commit +", commit.get(g + 1, 0) - c)

if __name__ == "__main__":
    main(*sys.argv[1:2])
