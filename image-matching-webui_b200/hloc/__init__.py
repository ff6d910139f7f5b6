"""Mirror of the accelerated part of `imcui.hloc` (reference: imcui/hloc/__init__.py)."""
import logging
from pathlib import Path

logger = logging.getLogger("hloc_b200")
MODEL_REPO_ID = "Realcat/imcui_checkpoints"  # hloc/__init__.py:66 (offline here: see utils/base_model.py)
WEIGHTS_DIR = Path(__file__).resolve().parent.parent.parent / "weights"
