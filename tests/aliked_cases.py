"""ALIKED golden cases, shared by tools/make_golden.py (generator) and the tests."""
ALIKED_CASES = {  # tag: (seed, H, W, rgb, conf)
    "s": (0, 240, 320, False, {"detection_threshold": 0.2, "max_num_keypoints": -1}),
    "m": (1, 480, 640, True, {"detection_threshold": 0.2, "max_num_keypoints": -1}),
    "cap": (1, 480, 640, True, {"detection_threshold": 0.2, "max_num_keypoints": 128}),
    "dense": (3, 480, 640, True, {"detection_threshold": 0.1, "max_num_keypoints": -1}),
    "densecap": (3, 480, 640, True, {"detection_threshold": 0.1, "max_num_keypoints": 1024}),
    "pad": (2, 250, 330, True, {"detection_threshold": 0.2, "max_num_keypoints": -1}),
    "topk": (0, 240, 320, False, {"detection_threshold": -1, "max_num_keypoints": 300}),
    "mean": (0, 240, 320, False, {"detection_threshold": 0.99, "max_num_keypoints": 512}),
}
