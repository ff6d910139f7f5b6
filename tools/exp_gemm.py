"""Experiment: split-fp16 GEMM timings with the splitter / epilogue ablations (imw_debug_set_gemm_ablate)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import imcui_b200
from imcui_b200 import _lib as L, ops
dev = torch.device("cuda:0")
lib = L.lib()
torch.manual_seed(0)

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

M = 131072
print("== timing (M = 131072 rows: 64 pairs x 2 x 1024 keypoints): full | no split | no epilogue | neither", flush=True)
for (N, K) in ((768, 256), (256, 256), (512, 512), (256, 512), (512, 256)):
    A = torch.randn(M, K, device=dev); Wt = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    Wp = ops.with_f16_planes(Wt).contiguous()
    out = torch.empty(M, N, device=dev)
    args = (L.ptr(A), L.ptr(Wp), L.ptr(b), L.ptr(out), M, N, K)
    r = {}
    for mode in (0, 2, 4, 6):
        lib.imw_debug_set_gemm_ablate(mode)
        r[mode] = timeit(lambda: L.check(lib.imw_debug_gemm_tf32(*args, 5, L.stream_ptr(dev))))
    fl = 2.0 * M * N * K / 1e9
    print(f"  N={N} K={K}: {r[0]:.3f} ms ({fl / r[0]:.0f} TFLOP/s) | {r[2]:.3f} | {r[4]:.3f} | {r[6]:.3f}", flush=True)
lib.imw_debug_set_gemm_ablate(0)
