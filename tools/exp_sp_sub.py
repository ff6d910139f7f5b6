"""Experiment: SuperPoint conv-stack sub-batch size vs step time (bench workload)."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import imcui_b200
from imcui_b200 import _lib as L, engine
from imcui_b200.utils import synth
dev = torch.device("cuda:0")
P = 64
a, b = synth.make_pair_batch(list(range(P)))
imgs_u8 = torch.from_numpy(np.stack([a, b], 1).reshape(2 * P, 480, 640)).to(dev)
lib = L.lib()
lib.imw_debug_set_sp_sub.restype = __import__("ctypes").c_int
for sub in (8, 16, 32, 64, 128):
    lib.imw_debug_set_sp_sub(sub)
    L.workspaces._bufs.clear() if hasattr(L.workspaces, "_bufs") else None
    eng = engine.PairEngine(dev, P)
    imgs = eng.to_float(imgs_u8)
    for _ in range(2):
        eng.match_device(imgs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        sp, lg = eng.match_device(imgs)
    e1.record(); torch.cuda.synchronize()
    print(f"SP_SUB {sub:4d}: {e0.elapsed_time(e1) / 4:.2f} ms/step, matches {(lg['matches'][0::2] > -1).sum().item()}", flush=True)
