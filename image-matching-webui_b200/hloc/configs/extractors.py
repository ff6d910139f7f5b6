"""Extractor confs of the accelerated path (values from the reference's hloc/configs/extractors.py:9-52,231-258)."""
_SP_PRE = {"grayscale": True, "force_resize": True, "resize_max": 1600, "width": 640, "height": 480, "dfactor": 8}

confs = {
    "superpoint_aachen": {
        "output": "feats-superpoint-n4096-r1024",
        "model": {"name": "superpoint", "nms_radius": 3, "max_keypoints": 4096, "keypoint_threshold": 0.005},
        "preprocessing": dict(_SP_PRE),
    },
    "superpoint_max": {
        "output": "feats-superpoint-n4096-rmax1600",
        "model": {"name": "superpoint", "nms_radius": 3, "max_keypoints": 4096, "keypoint_threshold": 0.005},
        "preprocessing": dict(_SP_PRE),
    },
    "superpoint_inloc": {
        "output": "feats-superpoint-n4096-r1600",
        "model": {"name": "superpoint", "nms_radius": 4, "max_keypoints": 4096, "keypoint_threshold": 0.005},
        "preprocessing": {"grayscale": True, "resize_max": 1600},
    },
    "aliked-n16": {  # extractors.py:245-258
        "output": "feats-aliked-n16",
        "model": {"name": "aliked", "model_name": "aliked-n16", "max_num_keypoints": -1, "detection_threshold": 0.2, "nms_radius": 2},
        "preprocessing": {"grayscale": False, "resize_max": 1024},
    },
    "aliked-n16-rot": {  # extractors.py:231-244
        "output": "feats-aliked-n16-rot",
        "model": {"name": "aliked", "model_name": "aliked-n16rot", "max_num_keypoints": -1, "detection_threshold": 0.2, "nms_radius": 2},
        "preprocessing": {"grayscale": False, "resize_max": 1024},
    },
}
