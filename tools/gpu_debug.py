"""Stage-by-stage diagnostics on the GPU box (writes gpurun_out/debug_*.txt)."""
import sys, time, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
torch.set_grad_enabled(False)
import oracle
from oracle import superpoint as osp, lightglue as olg
from imcui_b200 import ops
from imcui_b200.hloc import extractors, matchers
from imcui_b200.hloc.utils.base_model import dynamic_load

dev = torch.device("cuda:0")
G = ROOT / "tests" / "golden"
out = []
def log(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); out.append(s)

g = dict(np.load(G / "sp_real.npz"))
images = torch.from_numpy(g["images"])
w = oracle.load_weights("superpoint_v1.pt")
conf = {"nms_radius": 3, "max_keypoints": 1024, "keypoint_threshold": 0.015, "remove_borders": 4}
sp = dynamic_load(extractors, "superpoint")(conf).eval().to(dev)
o = ops.superpoint_forward(sp._bufs(), images.to(dev), sp.conf, 1024, want_dense=True)
torch.cuda.synchronize()
ref = osp.forward(w, images, conf, return_dense=True)
for b in range(2):
    d = (o["dense_scores"][b].cpu() - ref["dense_scores"][b]).abs()
    log("dense scores img", b, "max abs diff", float(d.max()), "mean", float(d.mean()))
    n = int(o["counts"][0, b]); nt = int(o["counts"][1, b])
    k = o["keypoints"][b, :n].cpu(); rk = ref["keypoints"][b]
    a, bb = {tuple(x) for x in k.tolist()}, {tuple(x) for x in rk.tolist()}
    log("kpts img", b, "n", n, "total", nt, "ref", len(rk), "lost", len(bb - a), "gained", len(a - bb), "order equal", torch.equal(k, rk) if k.shape == rk.shape else False)
    if k.shape == rk.shape and torch.equal(k, rk):
        log("  score diff", float((o["scores"][b, :n].cpu() - ref["scores"][b]).abs().max()),
            "desc diff", float((o["descriptors"][b, :n].cpu().t() - ref["descriptors"][b]).abs().max()))
# timing SuperPoint batch
for B in (2, 16):
    x = images[:1].repeat(B, 1, 1, 1).to(dev)
    for _ in range(2): ops.superpoint_forward(sp._bufs(), x, sp.conf, 1024)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(3): ops.superpoint_forward(sp._bufs(), x, sp.conf, 1024)
    torch.cuda.synchronize(); log("SP batch", B, "ms/img", (time.time() - t) / 3 / B * 1e3)

# LightGlue
from conftest import lg_pair_from_source
gl = dict(np.load(G / "lg_synth.npz")); cache = {}
def golden(name):
    if name not in cache: cache[name] = dict(np.load(G / f"{name}.npz"))
    return cache[name]
wl = oracle.load_weights("superpoint_lightglue.pt")
lg = dynamic_load(matchers, "lightglue")({"match_threshold": 0.2, "depth_confidence": -1, "width_confidence": -1}).eval().to(dev)
k0, d0, k1, d1 = lg_pair_from_source(golden, gl["sources"][0])
t = lambda a: torch.from_numpy(a).to(dev)
trace = []
ro = olg.forward(wl, torch.from_numpy(k0)[None], torch.from_numpy(d0).t().contiguous()[None], torch.from_numpy(k1)[None],
                 torch.from_numpy(d1).t().contiguous()[None], {"depth_confidence": -1, "width_confidence": -1}, trace=trace)
data = {"image0": None, "image1": None, "keypoints0": t(k0)[None], "keypoints1": t(k1)[None], "scores0": None, "scores1": None,
        "descriptors0": t(d0)[None], "descriptors1": t(d1)[None]}
po = lg(data)
torch.cuda.synchronize()
m0 = po["matches0"][0].cpu()
log("LG full: stop", po["stop"], ro["stop"], "matches", int((m0 > -1).sum()), int((ro["matches0"] > -1).sum()),
    "equal", torch.equal(m0, ro["matches0"][0]), "score diff", float((po["matching_scores0"][0].cpu() - ro["matching_scores0"][0]).abs().max()))
for P in (1, 16):
    kp = torch.zeros(2 * P, 1024, 2, device=dev); ds = torch.zeros(2 * P, 1024, 256, device=dev)
    for p in range(P):
        kp[2 * p] = t(k0); kp[2 * p + 1] = t(k1); ds[2 * p] = t(d0).t(); ds[2 * p + 1] = t(d1).t()
    counts = torch.full((2 * P,), 1024, dtype=torch.int32, device=dev)
    for mode, kc in (("full", {"depth_confidence": -1, "width_confidence": -1, "filter_threshold": 0.2, "pruning_min_kpts": 1536}),
                     ("adaptive", {"depth_confidence": 0.95, "width_confidence": 0.99, "filter_threshold": 0.2, "pruning_min_kpts": 1536})):
        for _ in range(2): ops.lightglue_forward(lg._bufs(), 9, kp, ds, counts, kc)
        torch.cuda.synchronize(); tt = time.time()
        for _ in range(3): r = ops.lightglue_forward(lg._bufs(), 9, kp, ds, counts, kc)
        torch.cuda.synchronize(); log("LG", mode, "pairs", P, "ms/pair", (time.time() - tt) / 3 / P * 1e3, "stop", r["stop"][:4].tolist())
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "debug_stage.txt").write_text("\n".join(out))
