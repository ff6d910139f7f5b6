"""CPU restatement of the dense-match aggregation of imcui/hloc/match_dense.py:37-121 and :299-404 -- TEST INFRASTRUCTURE
(see oracle/__init__.py).  Plain NumPy / dict / Counter, the reference's own data model; pinned against outputs of the
UNMODIFIED reference module (tests/golden/dense_agg.npz, minted by tools/make_golden.py dense_agg)."""
from collections import Counter, defaultdict
from itertools import chain

import numpy as np


def to_cpts(kpts, ps):
    """:37-40"""
    if ps > 0.0:
        kpts = np.round(np.round((kpts + 0.5) / ps) * ps - 0.5, 2)
    return [tuple(c) for c in kpts]


def assign_keypoints(kpts, other_cpts, max_error, update=False, ref_bins=None, scores=None, cell_size=None):
    """:43-83 (nearest keypoint by brute force instead of scipy's KDTree: same arg-min up to exact distance ties)"""
    if not update:
        if len(other_cpts) == 0 or len(kpts) == 0:
            return np.full(len(kpts), -1)
        o = np.asarray(other_cpts, np.float64)
        d = np.sqrt(((np.asarray(kpts, np.float64)[:, None] - o[None]) ** 2).sum(-1))
        ids = d.argmin(1)
        ids[d[np.arange(len(ids)), ids] > max_error] = -1
        return ids
    ps = max(cell_size if cell_size is not None else max_error, max_error)
    cpts, bpts = to_cpts(kpts, ps), to_cpts(kpts, int(max_error))
    cp_to_id = {v: i for i, v in enumerate(other_cpts)}
    out = []
    for i, (c, b) in enumerate(zip(cpts, bpts)):
        if c not in cp_to_id:
            cp_to_id[c] = len(cp_to_id)
            other_cpts.append(c)
            if ref_bins is not None:
                ref_bins.append(Counter())
        if ref_bins is not None:
            ref_bins[cp_to_id[c]][b] += scores[i] if scores is not None else 1
        out.append(cp_to_id[c])
    return np.array(out)


def kpids_to_matches0(ids0, ids1, scores):
    """:86-121: keep, per id on either side, the best-scoring match; intersection; scatter"""
    valid = (ids0 != -1) & (ids1 != -1)
    m = np.stack([ids0[valid], ids1[valid]], -1).reshape(-1, 2)
    sc = scores[valid]
    if len(m) == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.float16)

    def best_per_value(col):
        best = {}
        for i, v in enumerate(col):
            if v not in best or sc[i] > sc[best[v]]:
                best[v] = i
        return set(best.values())
    keep = sorted(best_per_value(m[:, 0]) & best_per_value(m[:, 1]))
    m, sc = m[keep], sc[keep]
    if len(m) == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.float16)
    n = m[:, 0].max() + 1
    m0, s0 = -np.ones(n), np.zeros(n)
    m0[m[:, 0]], s0[m[:, 0]] = m[:, 1], sc
    return m0.astype(np.int32), s0.astype(np.float16)


def aggregate_matches(conf, pairs, inputs, required_queries=None, max_kps=None, cpdict=None):
    """:299-404 over in-memory arrays: inputs {pair: (kpts0, kpts1, scores)} -> ({pair: (matches0, scores0)}, {name: (kpts, score)})"""
    cpdict = defaultdict(list, cpdict or {})
    bindict = defaultdict(list)
    if required_queries is None:
        required_queries = set(sum(pairs, ()))
    required_queries = set(required_queries) - {k for k, v in cpdict.items() if isinstance(v, np.ndarray)}
    per_q = Counter(list(chain(*pairs)))
    pairs = [p for _, p in sorted(zip([min(per_q[i], per_q[j]) for i, j in pairs], pairs))]
    out_m, out_f = {}, {}
    for n0, n1 in pairs:
        k0, k1, sc = inputs[(n0, n1)]
        u0, u1 = n0 in required_queries, n1 in required_queries
        e0, c0 = (0.0, 0.0) if (u0 and not u1 and max_kps is None) else (conf["max_error"], conf["cell_size"])
        i0 = assign_keypoints(k0, cpdict[n0], e0, u0, bindict[n0], sc, c0)
        i1 = assign_keypoints(k1, cpdict[n1], conf["max_error"], u1, bindict[n1], sc, conf["cell_size"])
        out_m[(n0, n1)] = kpids_to_matches0(i0, i1, sc)
        for n in (n0, n1):
            per_q[n] -= 1
            if per_q[n] > 0 or n not in required_queries:
                continue
            score = [c.most_common(1)[0][1] for c in bindict[n]]
            kp = np.array([c.most_common(1)[0][0] for c in bindict[n]], dtype=np.float32)
            if max_kps:
                top = np.argsort(score)[::-1][: min(max_kps, len(kp))]
                kp, score = kp[top], np.array(score)[top]
            cpdict[n] = kp
            out_f[n] = (kp, np.asarray(score))
            del bindict[n]
    return out_m, out_f
