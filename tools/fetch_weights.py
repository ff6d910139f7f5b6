"""Convert the model parameters found in the reference checkout into flat tensor dicts
under weights/ (git-ignored, shipped to the GPU box by gpurun like the built .so).

The reference downloads these at run time from the HF hub (hloc/utils/base_model.py:37-43,
hloc/extractors/superpoint.py:45-49, hloc/matchers/lightglue.py:38-49); there is no network
here, so SURVEY.md 8(c)'s in-tree copies are used:
  superpoint_v1.pth                      -> weights/superpoint_v1.pt
  gim_lightglue_100h.ckpt (model.*)      -> weights/superpoint_lightglue.pt   (GIM-trained SP+LG)
  superglue_outdoor.pth                  -> weights/superglue_outdoor.pt   (hloc conf default: weights=outdoor)
Only parameters (data) are converted; no reference source code is copied.
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).parent))
import ref_import as R  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "weights"


def main(force=False):
    if not R.available():
        print("[fetch_weights] reference checkout not present; nothing to do")
        return False
    OUT.mkdir(exist_ok=True)
    jobs = {
        "superpoint_v1.pt": lambda: torch.load(str(R.SP_WEIGHTS), map_location="cpu"),
        "superpoint_lightglue.pt": R.lightglue_state_dict,
        "superglue_outdoor.pt": lambda: torch.load(str(R.SG_WEIGHTS / "superglue_outdoor.pth"), map_location="cpu"),
    }
    for name, fn in jobs.items():
        dst = OUT / name
        if dst.exists() and not force:
            continue
        sd = {k: v.detach().float().contiguous().clone() for k, v in fn().items() if torch.is_tensor(v)}
        torch.save(sd, str(dst))
        print(f"[fetch_weights] wrote {dst} ({len(sd)} tensors)")
    return True


if __name__ == "__main__":
    main(force="--force" in sys.argv)
