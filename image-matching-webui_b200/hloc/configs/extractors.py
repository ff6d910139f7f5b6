"""Extractor confs of the accelerated path (values from the reference's hloc/configs/extractors.py:9-52)."""
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
}
