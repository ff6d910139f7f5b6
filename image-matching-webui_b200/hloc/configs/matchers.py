"""Matcher confs of the accelerated path (values from the reference's hloc/configs/matchers.py:34-50,217-232,249-287)."""
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
    "aliked-lightglue": {  # matchers.py:68-84
        "output": "matches-aliked-lightglue",
        "model": {"name": "lightglue", "match_threshold": 0.2, "width_confidence": 0.99, "depth_confidence": 0.95,
                  "features": "aliked", "model_name": "aliked_lightglue.pth"},
        "preprocessing": dict(_PRE),
    },
    "loftr": {  # matchers.py:249-267
        "output": "matches-loftr",
        "model": {"name": "loftr", "weights": "outdoor", "max_keypoints": 2000, "match_threshold": 0.2},
        "preprocessing": {"grayscale": True, "resize_max": 1024, "dfactor": 8, "width": 640, "height": 480, "force_resize": True},
        "max_error": 1, "cell_size": 1,
    },
    "minima_loftr": {  # matchers.py:268-287
        "output": "matches-minima_loftr",
        "model": {"name": "loftr", "weights": "outdoor", "model_name": "minima_loftr.ckpt", "max_keypoints": 2000, "match_threshold": 0.2},
        "preprocessing": {"grayscale": True, "resize_max": 1024, "dfactor": 8, "width": 640, "height": 480, "force_resize": False},
        "max_error": 1, "cell_size": 1,
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
