"""Dense-match aggregation -- the semi-dense -> sparse conversion of imcui/hloc/match_dense.py:37-121 (to_cpts,
assign_keypoints, get_unique_matches, matches_to_matches0, kpids_to_matches0) and :299-404 (aggregate_matches): the
detector-free matcher's per-pair keypoints are snapped to a grid, merged into ONE keypoint set per image across all of its
pairs (score-weighted voting of finer bins inside each cell), and every pair's correspondences are re-expressed as
`matches0` over those sets -- the files hloc's SfM / localisation pipelines read.

Division of labour.  Per pair, the array work runs on the GPU (csrc/dense_agg.cu: grid quantisation in the reference's fp32
arithmetic, nearest-keypoint assignment, n-to-1 conflict resolution with per-id atomic arg-max, matches0 scatter).  The
keypoint numbering of an image is first-come-first-numbered over the sorted pair list -- sequential by definition -- and is
kept on the host in `CellIndex`, vectorised with sorted key arrays instead of the reference's per-keypoint dict / Counter
loops: ids, vote sums (fp32, added in keypoint order like Counter's `+=`) and the tie rule of `Counter.most_common(1)`
(first inserted wins) are reproduced exactly.
"""
from collections import Counter
from itertools import chain

import numpy as np
import torch

from . import logger
from .. import ops
from .utils.parsers import names_to_pair
from .utils.store import open_store


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("dense_aggregate: the quantisation / matching kernels need a CUDA device (no CPU fallback)")
    return torch.device("cuda")


def to_cpts(kpts, ps):
    """match_dense.py:37-40 on the GPU.  kpts [K,2] -> (cells int32 [K,2], coords fp32 [K,2]); `coords` rows are the tuples the
    reference hashes, `cells` their integer grid indices (the raw fp32 bit patterns when ps == 0)."""
    k = np.ascontiguousarray(kpts, dtype=np.float32).reshape(-1, 2)
    if len(k) == 0:
        return np.zeros((0, 2), np.int32), np.zeros((0, 2), np.float32)
    cells, coords = ops.quantize_keypoints(torch.from_numpy(k).to(_dev())[None], float(ps))
    return cells[0].cpu().numpy(), coords[0].cpu().numpy()


def _pack(cells):
    """two int32 grid indices -> one int64 key (bijective)"""
    c = cells.astype(np.int64)
    return (c[:, 0] << 32) ^ (c[:, 1] & 0xFFFFFFFF)


def _mix(ids, cells):
    """(keypoint id, bin cell) -> one 64-bit key: splitmix64 of the bijective cell key, offset by the id.  Not bijective; two
    different (id, bin) pairs of one image collide with probability ~ n^2 / 2^65 (1e-8 for a million bins)."""
    with np.errstate(over="ignore"):
        z = _pack(cells).astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15) * (ids.astype(np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.astype(np.int64)


class _KeyTable:
    """int64 key -> dense id in first-appearance order (what `cp_to_id` / Counter insertion order give), vectorised: sorted
    key array + searchsorted for lookup, stable unique for the numbering of new keys."""

    def __init__(self):
        self.keys = np.zeros(0, np.int64)    # sorted
        self.ids = np.zeros(0, np.int64)
        self.n = 0

    def lookup_or_add(self, k):
        """k int64 [K] -> (ids [K], is_new [number of ids added], first index in k of every added id)"""
        pos = np.searchsorted(self.keys, k)
        hit = (pos < len(self.keys)) & (self.keys[np.minimum(pos, max(len(self.keys) - 1, 0))] == k) if len(self.keys) else np.zeros(len(k), bool)
        out = np.empty(len(k), np.int64)
        out[hit] = self.ids[pos[hit]]
        first_idx = np.zeros(0, np.int64)
        if (~hit).any():
            miss = np.nonzero(~hit)[0]
            uk, first = np.unique(k[miss], return_index=True)
            order = np.argsort(first, kind="stable")             # numbering follows the order of first appearance
            new_ids = np.empty(len(uk), np.int64)
            new_ids[order] = self.n + np.arange(len(uk))
            first_idx = miss[first[order]]
            out[miss] = new_ids[np.searchsorted(uk, k[miss])]
            self.keys = np.concatenate([self.keys, uk])
            self.ids = np.concatenate([self.ids, new_ids])
            srt = np.argsort(self.keys, kind="stable")
            self.keys, self.ids = self.keys[srt], self.ids[srt]
            self.n += len(uk)
        return out, first_idx


class CellIndex:
    """Growing keypoint set of one image (the reference's `cpdict[name]` list + `bindict[name]` list of Counters)."""

    def __init__(self):
        self.cells = _KeyTable()                   # cell key -> keypoint id
        self.coords = np.zeros((0, 2), np.float32)  # cell coordinate per id (what cpdict holds until the bins are resolved)
        self.bins = _KeyTable()                    # (id, bin key) -> vote slot
        self.bin_id = np.zeros(0, np.int64)
        self.bin_xy = np.zeros((0, 2), np.float32)
        self.bin_score = None

    def __len__(self):
        return self.cells.n

    def assign(self, kpts, max_error, cell_size=None, scores=None, vote=True):
        """assign_keypoints(update=True) (:60-83): quantise to cells of max(cell_size, max_error), number new cells, and add each
        keypoint's score to the bin (quantised at int(max_error)) it falls in.  Returns the keypoint ids [K]."""
        ps = max(cell_size if cell_size is not None else max_error, max_error)
        cells, coords = to_cpts(kpts, ps)
        ids, first = self.cells.lookup_or_add(_pack(cells))
        if len(first):
            self.coords = np.concatenate([self.coords, coords[first]])
        if vote and len(ids):
            bcell, bxy = to_cpts(kpts, int(max_error))
            slots, bfirst = self.bins.lookup_or_add(_mix(ids, bcell))
            w = np.ones(len(ids), np.int64) if scores is None else np.asarray(scores)
            if self.bin_score is None:
                self.bin_score = np.zeros(0, w.dtype)       # Counter sums in the dtype of the scores it is fed
            if len(bfirst):
                self.bin_id = np.concatenate([self.bin_id, ids[bfirst]])
                self.bin_xy = np.concatenate([self.bin_xy, bxy[bfirst]])
                self.bin_score = np.concatenate([self.bin_score, np.zeros(len(bfirst), w.dtype)])
            np.add.at(self.bin_score, slots, w)    # sequential adds in keypoint order, like Counter's `+=`
        return ids

    def resolve(self):
        """`c.most_common(1)[0]` per keypoint (:366-368): the best-voted bin's coordinate and its score; ties -> first inserted."""
        n = len(self)
        if self.bin_score is None:
            return self.coords.copy(), np.zeros(n, np.float32)
        order = np.lexsort((np.arange(len(self.bin_id)), -self.bin_score.astype(np.float64), self.bin_id))
        firsts = order[np.concatenate([[True], self.bin_id[order][1:] != self.bin_id[order][:-1]])] if len(order) else order
        kp = np.zeros((n, 2), np.float32); sc = np.zeros(n, self.bin_score.dtype)
        kp[self.bin_id[firsts]] = self.bin_xy[firsts]
        sc[self.bin_id[firsts]] = self.bin_score[firsts]
        return kp, sc


def assign_keypoints(kpts, other_cpts, max_error, update=False, ref_bins=None, scores=None, cell_size=None):
    """Signature of match_dense.py:43-83.  `other_cpts`: np.ndarray [M,2] of fixed keypoints (update=False: nearest-keypoint
    search on the GPU instead of a scipy KDTree) or a CellIndex (update=True; votes are recorded when ref_bins is not None)."""
    if not update:
        if len(other_cpts) == 0 or len(kpts) == 0:
            return np.full(len(kpts), -1)
        pts = other_cpts.coords if isinstance(other_cpts, CellIndex) else np.asarray(other_cpts, np.float32)
        d = _dev()
        ids = ops.nearest_point(torch.from_numpy(np.ascontiguousarray(kpts, dtype=np.float32)).to(d), torch.from_numpy(np.ascontiguousarray(pts)).to(d), max_error)
        return ids.cpu().numpy().astype(np.int64)
    assert isinstance(other_cpts, CellIndex), "update=True grows a CellIndex"
    return other_cpts.assign(kpts, max_error, cell_size, scores, vote=ref_bins is not None)


def kpids_to_matches0(kpt_ids0, kpt_ids1, scores):
    """match_dense.py:111-121 (+ get_unique_matches :94-108, matches_to_matches0 :99-108) on the GPU: -> (matches0 int32 [n_kps0],
    scores0 fp16 [n_kps0]) with n_kps0 = largest kept id0 + 1."""
    k0, k1 = np.asarray(kpt_ids0, np.int64), np.asarray(kpt_ids1, np.int64)
    if len(k0) == 0 or not ((k0 != -1) & (k1 != -1)).any():
        return np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float16)
    d = _dev()
    id_cap = int(max(k0.max(), k1.max())) + 1
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(d)[None]
    m0, s0, nk = ops.unique_matches(t(k0, np.int32), t(k1, np.int32), t(scores, np.float32), torch.tensor([len(k0)], dtype=torch.int32, device=d), id_cap)
    n = int(nk[0])
    return m0[0, :n].cpu().numpy(), s0[0, :n].cpu().numpy()


def aggregate_matches(conf, pairs, match_path, feature_path, required_queries=None, max_kps=None, cpdict=None, bindict=None):
    """match_dense.py:299-404.  conf: {"max_error", "cell_size"}; `match_path` holds, per `names_to_pair` group, keypoints0 /
    keypoints1 / scores of the dense matcher; adds matches0 / matching_scores0 to every group and writes each finished image's
    keypoints (+ `score`) to `feature_path`.  cpdict: {name: np.ndarray of fixed keypoints | CellIndex}.  Returns cpdict."""
    cpdict = {} if cpdict is None else cpdict
    if required_queries is None:
        required_queries = set(sum(pairs, ()))
        try:
            with open_store(feature_path, "r") as st:     # default: do not overwrite existing features
                required_queries -= set(st.groups())
        except FileNotFoundError:
            pass
    required_queries = set(required_queries) - {k for k, v in cpdict.items() if isinstance(v, np.ndarray)}
    pairs_per_q = Counter(list(chain(*pairs)))
    pairs = [p for _, p in sorted(zip([min(pairs_per_q[i], pairs_per_q[j]) for i, j in pairs], pairs))]   # reduced RAM (:318-321)
    if required_queries:
        logger.info(f"Aggregating keypoints for {len(required_queries)} images.")
    n_kps = 0
    mstore, fstore = open_store(match_path, "a"), open_store(feature_path, "a")
    try:
        for name0, name1 in pairs:
            pair = names_to_pair(name0, name1)
            kpts0, kpts1, scores = mstore.read(pair, "keypoints0"), mstore.read(pair, "keypoints1"), mstore.read(pair, "scores")
            assert kpts0.shape[0] == scores.shape[0]
            upd = [name0 in required_queries, name1 in required_queries]
            # in localization the query keypoints are not binned (assumes the query is name0, :341-346)
            err0, cs0 = (0.0, 0.0) if (upd[0] and not upd[1] and max_kps is None) else (conf["max_error"], conf["cell_size"])
            ids = []
            for name, k, u, err, cs in ((name0, kpts0, upd[0], err0, cs0), (name1, kpts1, upd[1], conf["max_error"], conf["cell_size"])):
                if u:
                    cpdict.setdefault(name, CellIndex())
                    ids.append(assign_keypoints(k, cpdict[name], err, True, True, scores, cs))
                else:
                    ids.append(assign_keypoints(k, cpdict.get(name, np.zeros((0, 2), np.float32)), err))
            m0, s0 = kpids_to_matches0(ids[0], ids[1], scores)
            mstore.write_group(pair, {"matches0": m0, "matching_scores0": s0}, replace=False)
            for name in (name0, name1):                                  # an image is finished after its last pair
                pairs_per_q[name] -= 1
                if pairs_per_q[name] > 0 or name not in required_queries:
                    continue
                kp, kp_score = cpdict[name].resolve()
                if max_kps:                                              # top-k keypoints by vote (:371-376)
                    top = np.argsort(kp_score)[::-1][: min(max_kps, len(kp))]
                    kp, kp_score = kp[top], kp_score[top]
                cpdict[name] = kp
                fstore.write_group(name, {"keypoints": kp, "score": kp_score})
                n_kps += len(kp)
    finally:
        mstore.close()
        fstore.close()
    if required_queries:
        logger.info(f"Finished assignment, found {round(n_kps / len(required_queries), 1)} keypoints/image (avg.), total {n_kps}.")
    return cpdict
