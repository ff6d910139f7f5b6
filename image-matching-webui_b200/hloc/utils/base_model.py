"""Plugin base class and discovery -- same contract as the reference's
imcui/hloc/utils/base_model.py:9-55 (BaseModel merges default_conf into a mutable self.conf, checks
required_inputs with an AssertionError, dispatches to _init/_forward; dynamic_load returns the single
BaseModel subclass defined in `<root>.<name>`)."""
import inspect
import sys
from abc import ABCMeta, abstractmethod
from copy import copy

from torch import nn

from .. import WEIGHTS_DIR

# file the reference asks the HF hub for  ->  converted in-tree parameters (tools/fetch_weights.py)
_LOCAL_WEIGHTS = {
    "superglue/superpoint_v1.pth": "superpoint_v1.pt",
    "lightglue/superpoint_lightglue.pth": "superpoint_lightglue.pt",
    "superglue/superglue_outdoor.pth": "superglue_outdoor.pt",
    "superglue/superglue_indoor.pth": "superglue_indoor.pt",
}


class BaseModel(nn.Module, metaclass=ABCMeta):
    default_conf = {}
    required_inputs = []

    def __init__(self, conf):
        """Perform some logic and call the _init method of the child model."""
        super().__init__()
        self.conf = conf = {**self.default_conf, **conf}
        self.required_inputs = copy(self.required_inputs)
        self._init(conf)
        sys.stdout.flush()

    def forward(self, data):
        """Check the data and call the _forward method of the child model."""
        for key in self.required_inputs:
            assert key in data, "Missing key {} in data".format(key)
        return self._forward(data)

    @abstractmethod
    def _init(self, conf):
        raise NotImplementedError

    @abstractmethod
    def _forward(self, data):
        raise NotImplementedError

    def _download_model(self, repo_id=None, filename=None, **kwargs):
        """Reference: hf_hub_download(repo_id, filename) (base_model.py:37-43).  This box has no
        network; the same parameters are resolved from `weights/` (converted from the reference
        checkout by tools/fetch_weights.py).  Falls back to the HF hub when available."""
        local = WEIGHTS_DIR / _LOCAL_WEIGHTS.get(filename, filename)
        if local.exists():
            return local
        try:
            from huggingface_hub import hf_hub_download
            return hf_hub_download(repo_type="model", repo_id=repo_id, filename=filename)
        except Exception as e:  # pragma: no cover
            raise FileNotFoundError(f"weights for {filename} not found under {WEIGHTS_DIR} and hub download failed: {e}")


def dynamic_load(root, model):
    module_path = f"{root.__name__}.{model}"
    module = __import__(module_path, fromlist=[""])
    classes = inspect.getmembers(module, inspect.isclass)
    classes = [c for c in classes if c[1].__module__ == module_path]
    classes = [c for c in classes if issubclass(c[1], BaseModel)]
    assert len(classes) == 1, classes
    return classes[0][1]
