"""Matcher confs of the accelerated path (values from the reference's hloc/configs/matchers.py:34-50,217-232)."""
_PRE = {"grayscale": True, "resize_max": 1024, "dfactor": 8, "force_resize": False}

confs = {
    "superglue": {
        "output": "matches-superglue",
        "model": {"name": "superglue", "weights": "outdoor", "sinkhorn_iterations": 50, "match_threshold": 0.2},
        "preprocessing": dict(_PRE),
    },
    "superglue-fast": {
        "output": "matches-superglue-it5",
        "model": {"name": "superglue", "weights": "outdoor", "sinkhorn_iterations": 5, "match_threshold": 0.2},
    },
    "superpoint-lightglue": {
        "output": "matches-lightglue",
        "model": {"name": "lightglue", "match_threshold": 0.2, "width_confidence": 0.99, "depth_confidence": 0.95,
                  "features": "superpoint", "model_name": "superpoint_lightglue.pth"},
        "preprocessing": dict(_PRE),
    },
    "NN-mutual": {
        "output": "matches-NN-mutual",
        "model": {"name": "nearest_neighbor", "do_mutual_check": True, "match_threshold": 0.2},
    },
    "Dual-Softmax": {
        "output": "matches-Dual-Softmax",
        "model": {"name": "dual_softmax", "match_threshold": 0.01, "inv_temperature": 20},
    },
}
