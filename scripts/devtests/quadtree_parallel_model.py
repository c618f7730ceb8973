#!/usr/bin/env python
"""CPU model of the PARALLEL formulation of Quadtree::split / nodes2kpoints (src/ORBExtractor.cc:126-192) used by the
quadtree kernel's fast path, checked against the oracle's sequential restatement on random and adversarial corner lists.

The reference pops the fullest node of a multimap<count, node, greater> until enough nodes exist.  A child never holds
more corners than its parent, so the pop sequence is the list of ALL tree nodes sorted by
    (count desc, parent's pop position, child index)
cut at the first prefix whose running node count reaches `need`.  Unrolled, "parent's pop position" is the chain of
ancestor counts, so two nodes of equal count compare by (count of parent, count of grandparent, ..., +inf for the root)
descending and finally by their path.  With every node's count known up front (corners sorted by their descent key: a
node is a contiguous run) the whole loop becomes: histogram of the node-count deltas -> the bucket c* in which the loop
stops -> rank the nodes of bucket c* -> popped set -> leaves -> drop the surplus (at most 3) from the end of the order.

The model bails out (returns None) exactly where the kernel falls back to the sequential loop: a node that would have to be
split below the key depth, or a pop whose delta is negative (all its corners on split lines) at or above c*.

    python scripts/devtests/quadtree_parallel_model.py [n_cases]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

KEY_LEVELS = 9
DROP = 7
FIX = 40
INF = 1 << 30


def strip_bounds(w, h):
    n_ini = int(round(w / h))  # python round == C round for non-.5; .5 cases: C rounds half away, python half even
    import math

    n_ini = int(math.floor(w / h + 0.5))
    if n_ini < 1:
        return n_ini, []
    hx = np.float32(np.float64(w) / np.float64(n_ini))
    cols = [0]
    for k in range(1, n_ini):
        v = np.float32(np.float32(k) * hx)
        cols.append(int(round(float(v) * (1 << FIX))))
    cols.append(w << FIX)
    return n_ini, cols


def make_key(x, y, cols, K, roi_h):
    """qt_make_key: (strip, digits[1..9]) with DROP at the first depth where the corner sits on a midline; strip -1 = in no strip"""
    X = x << FIX
    strip = -1
    if 0 < y < roi_h:
        for k in range(K):
            if cols[k] < X < cols[k + 1]:
                strip = k
                break
    if strip < 0:
        return None
    ux, wx = (X - cols[strip]) << KEY_LEVELS, cols[strip + 1] - cols[strip]
    qx, x_exact = ux // wx, (ux % wx) == 0
    uy = y << KEY_LEVELS
    qy, y_exact = uy // roi_h, (uy % roi_h) == 0

    def first_line(q, exact):
        if not exact:
            return KEY_LEVELS + 1
        tz = (q & -q).bit_length() - 1 if q else KEY_LEVELS + 1
        return KEY_LEVELS - tz

    jd = max(1, min(first_line(qx, x_exact), first_line(qy, y_exact)))
    digits = []
    for j in range(1, KEY_LEVELS + 1):
        d = (((qy >> (KEY_LEVELS - j)) & 1) << 1) | ((qx >> (KEY_LEVELS - j)) & 1)
        digits.append(d if j < jd else (DROP if j == jd else 0))
    return (strip, *digits)


def parallel_quadtree(roi_w, roi_h, xs, ys, resp, need, stats=None):
    """-> sorted list of selected corner indices, or None where the kernel falls back to the sequential loop"""
    n = len(xs)
    def bail(why):
        if stats is not None:
            stats["bail"] = why
        return None

    if need < 2:
        return bail("need<2")
    K, cols = strip_bounds(roi_w, roi_h)
    if K < 1 or K > 31:
        return bail("strips")
    keyed = []
    for i in range(n):
        k = make_key(int(xs[i]), int(ys[i]), cols, K, roi_h)
        if k is not None:
            keyed.append((k, i))
    keyed.sort(key=lambda t: t[0])  # order inside equal keys is irrelevant
    m = len(keyed)
    if m == 0:
        return []
    skey = [t[0] for t in keyed]
    ord_ = [t[1] for t in keyed]

    def lcp(a, b):
        c = 0
        while c < KEY_LEVELS + 1 and a[c] == b[c]:
            c += 1
        return c

    L = [0] * (m + 1)
    for p in range(1, m):
        L[p] = lcp(skey[p - 1], skey[p])
    vd = []
    for k in skey:
        d = KEY_LEVELS
        for j in range(1, KEY_LEVELS + 1):
            if k[j] == DROP:
                d = j - 1
                break
        vd.append(d)

    def run_end(pos, d):
        e = pos + 1
        while e < m and skey[e][: d + 1] == skey[pos][: d + 1]:
            e += 1
        return e

    def children(pos, e, d):
        """non-empty children (digit, start, end) of the depth-d node [pos, e)"""
        out = []
        p = pos
        while p < e:
            dg = skey[p][d + 1]
            q = p + 1
            while q < e and skey[q][d + 1] == dg:
                q += 1
            if dg != DROP:
                out.append((dg, p, q))
            p = q
        return out

    # ---- nodes with >= 2 corners: (pos, depth) for depth in [L[pos], min(L[pos+1] - 1, vd[pos])]
    nodes = []  # dict per node
    deep_multi_max = 0  # largest count of a node AT the key depth with >= 2 corners (cannot be split by keys)
    neg_max = 0         # largest count of a node whose delta is negative
    for pos in range(m):
        for d in range(L[pos], min(L[pos + 1] - 1, vd[pos]) + 1):
            e = run_end(pos, d)
            cnt = e - pos
            assert cnt >= 2
            if d == KEY_LEVELS:
                deep_multi_max = max(deep_multi_max, cnt)
                nodes.append(dict(pos=pos, end=e, depth=d, cnt=cnt, delta=None))
                continue
            ch = children(pos, e, d)
            delta = len(ch) - 1
            if delta < 0:
                neg_max = max(neg_max, cnt)
            nodes.append(dict(pos=pos, end=e, depth=d, cnt=cnt, delta=delta))
    # strips: children of the root
    strips = []
    p = 0
    while p < m:
        q = p + 1
        while q < m and skey[q][0] == skey[p][0]:
            q += 1
        strips.append((skey[p][0], p, q))
        p = q
    live0 = len(strips)

    # ---- the bucket in which the loop stops
    hist = {}
    for nd in nodes:
        if nd["delta"] is not None:
            hist[nd["cnt"]] = hist.get(nd["cnt"], 0) + nd["delta"]
    c_star = None
    before = live0  # node count after all buckets above c_star
    if live0 < need:
        run = live0
        for c in sorted(set(nd["cnt"] for nd in nodes), reverse=True):
            if c <= deep_multi_max or c <= neg_max:
                return bail("deep" if c <= deep_multi_max else "negative")  # the loop would pop a node the keys cannot split / a negative delta: sequential fallback
            after = run + hist.get(c, 0)
            if after >= need:
                c_star, before = c, run
                break
            run = after
        if c_star is None:
            return []  # starved: the fullest node holds one corner while short of the quota -> the multimap drains (0 keypoints)

    # ---- materialise the nodes that can pop (count >= c*), with parent links, and rank them in pop order
    S = [nd for nd in nodes if c_star is not None and nd["cnt"] >= c_star]
    by_key = {(nd["pos"], nd["depth"]): i for i, nd in enumerate(S)}
    for nd in S:
        d, pos = nd["depth"], nd["pos"]
        if d == 0:
            nd["parent"], nd["child"] = -1, skey[pos][0]
        else:
            pp = pos
            while pp > 0 and skey[pp - 1][:d] == skey[pos][:d]:
                pp -= 1
            nd["parent"], nd["child"] = by_key[(pp, d - 1)], skey[pos][d]

    def before_in_pop_order(u, v):
        a, b = u, v
        pa = pb = None
        while True:
            if a == b:
                return (S[pa]["child"] if pa is not None else 0) < (S[pb]["child"] if pb is not None else 0)
            ca = INF if a < 0 else S[a]["cnt"]
            cb = INF if b < 0 else S[b]["cnt"]
            if ca != cb:
                return ca > cb
            pa, pb = a, b
            a = -1 if a < 0 else S[a]["parent"]
            b = -1 if b < 0 else S[b]["parent"]

    for i in range(len(S)):
        S[i]["rank"] = sum(1 for j in range(len(S)) if j != i and before_in_pop_order(j, i))
    popped = set()
    last_rank = -1
    if c_star is not None:
        order = sorted(range(len(S)), key=lambda i: S[i]["rank"])
        run = None
        for i in order:
            nd = S[i]
            if nd["cnt"] > c_star:
                popped.add(i)
                continue
            if run is None:
                run = before
            if run >= need:
                break
            popped.add(i)
            last_rank = nd["rank"]
            run += nd["delta"]
        assert run is not None and run >= need
    if stats is not None:
        stats["pops"] = len(popped) + 1
        stats["S"] = len(S)
        stats["nodes"] = len(nodes)
        stats["bucket"] = sum(1 for nd in S if nd["cnt"] == c_star) if c_star else 0

    # ---- leaves: non-empty children of popped nodes (and of the root) that were not popped themselves
    leaves = []  # (count, parent_rank, child idx, start, end)
    for (sid, p, q) in strips:
        i = by_key.get((p, 0))
        if i is None or i not in popped:
            leaves.append((q - p, -1, sid, p, q))
    for i in popped:
        nd = S[i]
        for (dg, p, q) in children(nd["pos"], nd["end"], nd["depth"]):
            j = by_key.get((p, nd["depth"] + 1))
            if j is None or j not in popped:
                leaves.append((q - p, nd["rank"], dg, p, q))
    total = len(leaves)
    take = min(need, total)
    # first `take` leaves in (count desc, insertion order): drop the surplus from the end of that order
    leaves.sort(key=lambda t: (-t[0], t[1], t[2]))
    sel = set()
    for (cnt, _, _, p, q) in leaves[:take]:
        best, best_i = 0, 0
        for pos in range(p, q):
            idx = ord_[pos]
            r = int(resp[idx])
            if r > best or (r == best and r > 0 and idx < best_i):
                best, best_i = r, idx
        sel.add(best_i)
    return sorted(sel)


def _unique_points(rng, w, h, n):
    n = min(n, (w - 6) * (h - 6))
    flat = rng.choice((w - 6) * (h - 6), n, replace=False)
    flat.sort()
    return (flat % (w - 6) + 3).astype(np.int32), (flat // (w - 6) + 3).astype(np.int32)


def main():
    from oracle import oracle_py as O

    O.build()
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(1)
    shapes = [(1209, 344), (608, 448), (314, 73), (100, 400), (64, 64), (1208, 344), (1888, 1048), (1000, 250), (513, 512)]
    ok = bail = 0
    reasons = {}
    worst = dict(pops=0, S=0, bucket=0)
    for case in range(n_cases):
        w, h = shapes[case % len(shapes)]
        need = int(rng.choice([2, 3, 7, 30, 60, 119, 300, 434, 900]))
        kind = case % 5
        n = int(rng.choice([0, 1, 2, 5, need - 1, need, need + 3, 2 * need, 3 * need, 5 * need, 2500]))
        n = max(0, min(n, w * h // 6))
        xs, ys = _unique_points(rng, w, h, n)
        if kind == 1 and n > 10:  # corners on strip bounds / midlines
            K, cols = strip_bounds(w, h)
            lines_x = [int(c >> FIX) for c in cols if (c >> FIX) << FIX == c] + [w // 2, w // 4, w // 8]
            lines_y = [h // 2, h // 4, 3 * h // 4, h // 8, 0, h]
            k = n // 3
            xs[:k] = rng.choice(lines_x, k)
            ys[k:2 * k] = rng.choice(lines_y, k)
        if kind == 2 and n > 10:  # tight clusters
            cx, cy = rng.integers(5, w - 8, max(1, n // 9)), rng.integers(5, h - 8, max(1, n // 9))
            pts = sorted({(int(a + dx), int(b + dy)) for a, b in zip(cx, cy) for dx in range(3) for dy in range(3)}, key=lambda p: (p[1], p[0]))
            xs, ys = np.array([p[0] for p in pts], np.int32), np.array([p[1] for p in pts], np.int32)
        key = xs.astype(np.int64) * 8192 + ys
        _, first = np.unique(key, return_index=True)
        first.sort()
        xs, ys = xs[first], ys[first]
        resp = rng.integers(6, 120, len(xs)).astype(np.int32) if kind != 3 else np.full(len(xs), 30, np.int32)
        exp, pops = O.quadtree_select(w, h, xs, ys, resp, need)
        st = {}
        got = parallel_quadtree(w, h, xs, ys, resp, need, st)
        if got is None:
            bail += 1
            reasons[(kind, st.get("bail"))] = reasons.get((kind, st.get("bail")), 0) + 1
            continue
        assert list(exp) == got, (case, w, h, len(xs), need, len(exp), len(got))
        if st:
            assert st["pops"] == pops, (st, pops)
            for k in worst:
                worst[k] = max(worst[k], st.get(k, 0))
        ok += 1
    print("fallback reasons (input kind, reason):", reasons)
    print(f"{ok} cases identical to the oracle, {bail} fell back to the sequential loop; worst {worst}")


if __name__ == "__main__":
    main()
