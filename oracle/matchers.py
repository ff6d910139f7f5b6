"""Oracle for the two first-party hloc matchers.
Test infrastructure only (see oracle/__init__.py)."""
import numpy as np
import torch


def find_nn(sim, ratio_thresh, distance_thresh):
    """hloc/matchers/nearest_neighbor.py:6-16."""
    sim_nn, ind_nn = sim.topk(2 if ratio_thresh else 1, dim=-1, largest=True)
    dist_nn = 2 * (1 - sim_nn)
    mask = torch.ones(ind_nn.shape[:-1], dtype=torch.bool)
    if ratio_thresh:
        mask = mask & (dist_nn[..., 0] <= (ratio_thresh ** 2) * dist_nn[..., 1])
    if distance_thresh:
        mask = mask & (dist_nn[..., 0] <= distance_thresh ** 2)
    matches = torch.where(mask, ind_nn[..., 0], ind_nn.new_tensor(-1))
    scores = torch.where(mask, (sim_nn[..., 0] + 1) / 2, sim_nn.new_tensor(0))
    return matches, scores


def mutual_check(m0, m1):
    """hloc/matchers/nearest_neighbor.py:19-24."""
    inds0 = torch.arange(m0.shape[-1])
    loop = torch.gather(m1, -1, torch.where(m0 > -1, m0, m0.new_tensor(0)))
    ok = (m0 > -1) & (inds0 == loop)
    return torch.where(ok, m0, m0.new_tensor(-1))


def nearest_neighbor(desc0, desc1, ratio_threshold=None, distance_threshold=None, do_mutual_check=True):
    """NearestNeighbor._forward, hloc/matchers/nearest_neighbor.py:38-66.  desc* [1,D,N]."""
    if desc0.size(-1) == 0 or desc1.size(-1) == 0:
        matches0 = torch.full(desc0.shape[:2], -1)
        return {"matches0": matches0, "matching_scores0": torch.zeros_like(matches0)}
    if desc0.size(-1) == 1 or desc1.size(-1) == 1:
        ratio_threshold = None
    sim = torch.einsum("bdn,bdm->bnm", desc0, desc1)
    matches0, scores0 = find_nn(sim, ratio_threshold, distance_threshold)
    if do_mutual_check:
        matches1, _ = find_nn(sim.transpose(1, 2), ratio_threshold, distance_threshold)
        matches0 = mutual_check(matches0, matches1)
    return {"matches0": matches0, "matching_scores0": scores0}


def dual_softmax(desc0, desc1, match_threshold=0.2, inv_temperature=20):
    """DualSoftMax._forward + dual_softmax_matcher, hloc/matchers/dual_softmax.py:8-71.
    desc* [1,D,N].  matching_scores0 is float64 and matches0 int64, as in the reference
    (NumPy scatter, last write wins, batch index ignored :31)."""
    if desc0.size(-1) == 0 or desc1.size(-1) == 0:
        matches0 = torch.full(desc0.shape[:2], -1)
        return {"matches0": matches0, "matching_scores0": torch.zeros_like(matches0)}
    B = desc0.shape[0]
    a = desc0 / desc0.norm(dim=1, keepdim=True)
    b = desc1 / desc1.norm(dim=1, keepdim=True)
    sim = torch.einsum("b c n, b c m -> b n m", a, b) * inv_temperature
    P = sim.softmax(dim=-2) * sim.softmax(dim=-1)
    mask = torch.nonzero(
        (P == P.max(dim=-1, keepdim=True).values)
        * (P == P.max(dim=-2, keepdim=True).values)
        * (P > match_threshold)
    ).numpy()
    matches0 = np.ones((B, P.shape[-2]), dtype=int) * (-1)
    scores0 = np.zeros((B, P.shape[-2]), dtype=float)
    matches0[:, mask[:, 1]] = mask[:, 2]
    scores0[:, mask[:, 1]] = P.numpy()[mask[:, 0], mask[:, 1], mask[:, 2]]
    return {"matches0": torch.from_numpy(matches0), "matching_scores0": torch.from_numpy(scores0)}
