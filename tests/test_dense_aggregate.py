"""Dense-match aggregation.  CPU: the oracle (oracle/dense.py) against golden outputs of the
UNMODIFIED reference module (tools/make_golden.py dense_agg).  GPU: the product path (hloc/dense_aggregate.py: GPU quantisation /
assignment / conflict resolution + vectorised host numbering) against the same goldens, through the group/dataset store."""
import numpy as np
import pytest

NAMES = ["q/a.jpg", "db/b.jpg", "db/c.jpg", "db/d.jpg"]
PAIRS = [(NAMES[i], NAMES[j]) for i in range(4) for j in range(i + 1, 4)]
CASES = {"sfm": ({"max_error": 1, "cell_size": 1}, None, False), "coarse": ({"max_error": 2, "cell_size": 8}, 300, False),
         "loc": ({"max_error": 4, "cell_size": 4}, None, True)}


def _key(a, b):
    return f"{a.replace('/', '-')}/{b.replace('/', '-')}"


def _case(g, case):
    conf, max_kps, fixed = CASES[case]
    pairs = [p for p in PAIRS if p[0] == NAMES[0]] if fixed else PAIRS
    inputs = {p: tuple(g[f"{case}/in/{_key(*p)}/{k}"] for k in ("keypoints0", "keypoints1", "scores")) for p in pairs}
    cp = {n: g[f"{case}/fixed/{n}"] for n in NAMES[1:]} if fixed else None
    return conf, max_kps, pairs, inputs, cp


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_equals_reference(golden, case):
    from oracle import dense as od
    g = golden("dense_agg")
    conf, max_kps, pairs, inputs, cp = _case(g, case)
    out_m, out_f = od.aggregate_matches(conf, list(pairs), inputs, required_queries={NAMES[0]} if cp else None, max_kps=max_kps, cpdict=cp)
    for p in pairs:
        assert np.array_equal(out_m[p][0], g[f"{case}/out/{_key(*p)}/matches0"]), (case, p)
        assert np.array_equal(out_m[p][1], g[f"{case}/out/{_key(*p)}/matching_scores0"]), (case, p)
    for n, (kp, sc) in out_f.items():
        assert np.array_equal(kp, g[f"{case}/feat/{n}/keypoints"]) and np.array_equal(sc, g[f"{case}/feat/{n}/score"]), (case, n)


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_aggregate_matches_equals_reference(golden, case, tmp_path):
    from imcui_b200.hloc import dense_aggregate as da
    from imcui_b200.hloc.utils.store import open_store
    g = golden("dense_agg")
    conf, max_kps, pairs, inputs, cp = _case(g, case)
    mpath, fpath = tmp_path / "matches.imw", tmp_path / "feats.imw"
    with open_store(mpath, "a") as st:
        for p, (k0, k1, sc) in inputs.items():
            st.write_group(_key(*p), {"keypoints0": k0, "keypoints1": k1, "scores": sc})
    da.aggregate_matches(conf, list(pairs), mpath, fpath, required_queries={NAMES[0]} if cp else None, max_kps=max_kps,
                         cpdict=dict(cp) if cp else None)
    with open_store(mpath, "r") as st:
        for p in pairs:
            m0, s0 = st.read(_key(*p), "matches0"), st.read(_key(*p), "matching_scores0")
            rm, rs = g[f"{case}/out/{_key(*p)}/matches0"], g[f"{case}/out/{_key(*p)}/matching_scores0"]
            assert m0.dtype == np.int32 and s0.dtype == np.float16
            assert np.array_equal(m0, rm), (case, p, int((m0 != rm).sum()) if m0.shape == rm.shape else (m0.shape, rm.shape))
            assert np.array_equal(s0, rs), (case, p)
    with open_store(fpath, "r") as st:
        names = [n for n in NAMES if f"{case}/feat/{n}/keypoints" in g]
        assert sorted(st.groups()) == sorted(names)
        for n in names:
            assert np.array_equal(st.read(n, "keypoints"), g[f"{case}/feat/{n}/keypoints"]), (case, n)
            assert np.array_equal(st.read(n, "score"), g[f"{case}/feat/{n}/score"]), (case, n)


@pytest.mark.gpu
def test_dense_kernels_against_numpy():
    """to_cpts / nearest keypoint / conflict resolution kernels on random data against the oracle's statements."""
    import torch
    from oracle import dense as od
    from imcui_b200 import ops
    from imcui_b200.hloc import dense_aggregate as da
    rng = np.random.default_rng(3)
    k = rng.uniform(-5, 700, (4000, 2)).astype(np.float32)
    k[:50] = np.round(k[:50]) + 0.5                                # exact .5 ties of np.round (half to even)
    for ps in (1.0, 2.0, 8.0, 0.0, 3.0):
        cells, coords = da.to_cpts(k, ps)
        ref = np.array(od.to_cpts(k, ps), np.float32)
        assert np.array_equal(coords, ref), ps
    pts = rng.uniform(0, 300, (777, 2)).astype(np.float32); q = rng.uniform(0, 300, (1500, 2)).astype(np.float32)
    ids = da.assign_keypoints(q, pts, 4.0)
    ref = od.assign_keypoints(q, pts, 4.0)
    assert (ids == ref).mean() > 0.999                              # equal up to exact distance ties / fp32 vs fp64 distance
    i0 = rng.integers(-1, 400, 5000); i1 = rng.integers(-1, 380, 5000); sc = rng.uniform(0, 1, 5000).astype(np.float32)
    m0, s0 = da.kpids_to_matches0(i0, i1, sc)
    rm, rs = od.kpids_to_matches0(i0, i1, sc)
    assert np.array_equal(m0, rm) and np.array_equal(s0, rs)
