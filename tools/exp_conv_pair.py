"""Experiment / A-B check: Cin = Cout = 64 convs on CTA pairs (imw_debug_set_conv_pair) vs the single-CTA kernels."""
import ctypes as C, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import imcui_b200
from imcui_b200 import _lib as L, ops
dev = torch.device("cuda:0")
lib = L.lib()
lib.imw_debug_set_conv_pair.restype = C.c_int
lib.imw_debug_set_conv_pair.argtypes = [C.c_int]
torch.manual_seed(0)


def merge(planes):
    return planes[0].float() + planes[1].float() / 2048.0


def fused(img, w1a, b1a, w1b, b1b, pool):
    B, H, W = img.shape
    wp = ops.split_f16_planes(w1b.permute(0, 2, 1).contiguous()).to(dev)   # [2][9][Cout][Cin]
    out = torch.empty(2, B, H // 2 if pool else H, W // 2 if pool else W, 64, dtype=torch.float16, device=dev)
    L.check(lib.imw_debug_conv1ab_fused(L.ptr(img), L.ptr(w1a), L.ptr(b1a), L.ptr(wp), L.ptr(b1b), L.ptr(out), B, H, W, int(pool), L.stream_ptr(dev)))
    return out


def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


modes = [int(a) for a in sys.argv[1:]] or [1, 2, 0]
print("== plain c64 conv vs fp32 CUDA-core conv", flush=True)
for (B, H, W, pool) in ((1, 16, 32, False), (2, 48, 64, True), (1, 24, 48, True), (1, 40, 16, False), (3, 16, 8, False), (1, 480, 640, True)):
    x = torch.rand(B, H, W, 64, device=dev)
    w = torch.randn(9, 64, 64, device=dev) / 24.0
    b = torch.randn(64, device=dev) * 0.1
    ref = ops.debug_conv3x3(x, w, b, relu=True, pool=pool, tensor_cores=False)
    for mode in modes:
        lib.imw_debug_set_conv_pair(mode)
        out = ops.debug_conv3x3(x, w, b, relu=True, pool=pool, tensor_cores=True)
        torch.cuda.synchronize()
        print(f"  {B}x{H}x{W} pool={pool} mode={mode}: max err {float((out - ref).abs().max()):.3e} (ref max {float(ref.abs().max()):.2f})", flush=True)

print("== fused conv1a+conv1b vs torch fp64", flush=True)
for (B, H, W, pool) in ((1, 16, 16, False), (2, 48, 64, True), (1, 40, 24, True), (3, 120, 160, True)):
    img = torch.rand(B, H, W, device=dev)
    w1a = torch.randn(9, 64, device=dev) / 3.0; b1a = torch.randn(64, device=dev) * 0.1
    w1b = torch.randn(9, 64, 64, device=dev) / 24.0; b1b = torch.randn(64, device=dev) * 0.1
    y = torch.nn.functional.conv2d(img[:, None].double(), w1a.view(3, 3, 1, 64).permute(3, 2, 0, 1).double(), b1a.double(), padding=1).relu()
    y = torch.nn.functional.conv2d(y.float().double(), w1b.view(3, 3, 64, 64).permute(3, 2, 0, 1).double(), b1b.double(), padding=1).relu()
    if pool: y = torch.nn.functional.max_pool2d(y, 2)
    ref = y.permute(0, 2, 3, 1).float()
    for mode in modes:
        lib.imw_debug_set_conv_pair(mode)
        out = merge(fused(img, w1a, b1a, w1b, b1b, pool))
        torch.cuda.synchronize()
        print(f"  {B}x{H}x{W} pool={pool} mode={mode}: max err {float((out - ref).abs().max()):.3e} (ref max {float(ref.abs().max()):.2f})", flush=True)

print("== timing (32 images)", flush=True)
B = 32
img = torch.rand(B, 480, 640, device=dev)
w1a = torch.randn(9, 64, device=dev) / 3.0; b1a = torch.randn(64, device=dev) * 0.1
w1b = torch.randn(9, 64, 64, device=dev) / 24.0; b1b = torch.randn(64, device=dev) * 0.1
xp = torch.rand(2, B, 240, 320, 64, device=dev).half()
wp = ops.split_f16_planes(w1b.permute(0, 2, 1).contiguous()).to(dev)
outp = torch.empty(2, B, 240, 320, 64, dtype=torch.float16, device=dev)
for mode in [m for m in modes if m != 2]:
    lib.imw_debug_set_conv_pair(mode)
    t_f = timeit(lambda: fused(img, w1a, b1a, w1b, b1b, True))
    t_c = timeit(lambda: L.check(lib.imw_debug_conv3x3_tc_planes(L.ptr(xp), L.ptr(wp), L.ptr(b1b), L.ptr(outp), B, 240, 320, 64, 64, 1, 0, L.stream_ptr(dev))))
    print(f"  mode={mode}: fused 480x640 {t_f:.3f} ms ({0.736 * 32 / 32 / t_f * 1e3 / 1e3 * 1e0:.1f}?), c64 240x320 {t_c:.3f} ms", flush=True)
    print(f"           fused {23.0 * B / t_f:.1f} TFLOP/s, c64 {2 * 9 * 64 * 64 * 240 * 320 * B / t_c / 1e9:.1f} TFLOP/s", flush=True)
lib.imw_debug_set_conv_pair(1)
print("== halo kernels (32 images): pair vs single CTA", flush=True)
for (H, W, Cin, Cout, pool) in ((120, 160, 64, 128, 0), (120, 160, 128, 128, 1), (60, 80, 128, 128, 0), (60, 80, 128, 256, 0)):
    xh = torch.rand(2, B, H, W, Cin, device=dev).half()
    wh = ops.split_f16_planes(torch.randn(9, Cout, Cin) / (9 * Cin) ** 0.5).to(dev)
    bh = torch.randn(Cout, device=dev) * 0.1
    oh = torch.empty(2, B, H // 2 if pool else H, W // 2 if pool else W, Cout, dtype=torch.float16, device=dev)
    res = {}
    for mode in (1, 0):
        lib.imw_debug_set_conv_pair(mode)
        tt = timeit(lambda: L.check(lib.imw_debug_conv3x3_tc_planes(L.ptr(xh), L.ptr(wh), L.ptr(bh), L.ptr(oh), B, H, W, Cin, Cout, 1, pool, L.stream_ptr(dev))))
        res[mode] = (tt, oh.clone())
    same = bool(torch.equal(res[0][1], res[1][1]))
    fl = 2 * 9 * Cin * Cout * H * W * B / 1e9
    print(f"  {H}x{W} {Cin}->{Cout} pool={pool}: pair {res[1][0]:.3f} ms ({fl / res[1][0]:.0f} TFLOP/s), single {res[0][0]:.3f} ms ({fl / res[0][0]:.0f} TFLOP/s), outputs identical: {same}", flush=True)

print("== ablations (timing only; results are garbage): 256 no producer compute, 512 no epilogue math/stores, 1024 no MMAs", flush=True)
for flags in (0, 256, 512, 1024, 256 + 512, 256 + 1024, 512 + 1024, 256 + 512 + 1024):
    lib.imw_debug_set_conv_pair(1 + flags)
    t_f = timeit(lambda: fused(img, w1a, b1a, w1b, b1b, True))
    t_c = timeit(lambda: L.check(lib.imw_debug_conv3x3_tc_planes(L.ptr(xp), L.ptr(wp), L.ptr(b1b), L.ptr(outp), B, 240, 320, 64, 64, 1, 0, L.stream_ptr(dev))))
    print(f"  flags={flags:5d}: fused {t_f:.3f} ms, c64 {t_c:.3f} ms", flush=True)
lib.imw_debug_set_conv_pair(1)
