"""B200-native image-pair matching engine behind the reference's hloc plugin API.

Importable as `imcui_b200` (see imcui_b200/__init__.py at the repo root; this directory's name has a
hyphen).  Layout mirrors the reference's `imcui/hloc` tree for the accelerated path only:

    hloc/utils/base_model.py      BaseModel, dynamic_load          (hloc/utils/base_model.py)
    hloc/extractors/superpoint.py SuperPoint                       (hloc/extractors/superpoint.py)
    hloc/matchers/lightglue.py    LightGlue                        (hloc/matchers/lightglue.py)
    hloc/matchers/nearest_neighbor.py, dual_softmax.py             (same names in the reference)
    hloc/extract_features.py      extract()                        (hloc/extract_features.py:106-170)
    hloc/match_features.py        match_images()                   (hloc/match_features.py:204-275)
    engine.py                     batched pair pipeline (bench / stream driver)
    csrc/                         sm_100a CUDA kernels + the C ABI (include/imw_b200.h)
"""
__version__ = "0.1.0"
