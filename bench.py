#!/usr/bin/env python
"""bench.py -- BASELINE metric: image-pairs/sec, SuperPoint+LightGlue, batch=64 synthetic 640x480 pairs.

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA engine through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   reference arm: the CPU path (oracle port of the
                                                            reference modules) on the host cores, rank 0 only
  torchrun ... bench.py --gpus N ...                        one rank per GPU, pairs sharded (weak scaling)

A step = one batch of 64 pairs per GPU through SuperPoint (x2 images) + LightGlue.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "image-pairs/sec @640x480 SuperPoint+LightGlue"
PAIRS_PER_GPU = 64
H, W = 480, 640
SP_CONF = {"nms_radius": 3, "keypoint_threshold": 0.005, "max_keypoints": 1024, "remove_borders": 4}
# algorithmic work (SURVEY.md 8(d) / BASELINE.md section 4)
SP_GFLOP_PER_IMAGE = 52.10
CONV1B_GFLOP_PER_IMAGE = 22.65


def lg_gflop(n, layers):
    return (layers * (4.98e6 * n + 3584.0 * n * n) + 2.6e5 * n + 512.0 * n * n) / 1e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fp32", action="store_true", help="everything on the fp32 CUDA-core path (no tensor cores)")
    ap.add_argument("--tf32", action="store_true", help="LightGlue linears as single TF32 (fast mode, not parity-grade)")
    ap.add_argument("--sp-simt", action="store_true", help="SuperPoint convs on fp32 CUDA cores")
    return ap.parse_args()


def synth_pairs(n_pairs, first_seed=0):
    from imcui_b200.utils import synth
    a, b = synth.make_pair_batch(range(first_seed, first_seed + n_pairs), H, W)
    out = np.empty((2 * n_pairs, H, W), np.uint8)
    out[0::2], out[1::2] = a, b
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Usable host cores: CPU affinity capped by the cgroup quota (os.cpu_count() reports the whole box)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        pass
    return n


def best_cpu_threads(images_u8):
    """Thread count that runs the reference CPU path fastest on this box (torch CPU convs stop scaling, and
    oversubscription past the cgroup quota is catastrophic): quick calibration on one SuperPoint image."""
    import oracle
    from oracle import superpoint as osp
    ws = oracle.load_weights("superpoint_v1.pt")
    img = torch.from_numpy(images_u8[:1].astype(np.float64) / 255.0).float()[:, None]
    cap = host_threads()
    best = (None, 1)
    for t in sorted({min(cap, c) for c in (8, 16, 32, 64, cap)}):
        torch.set_num_threads(t)
        osp.forward(ws, img, SP_CONF)
        t0 = time.perf_counter()
        osp.forward(ws, img, SP_CONF)
        dt = time.perf_counter() - t0
        if best[0] is None or dt < best[0]:
            best = (dt, t)
    return best[1]


def cpu_reference_pairs_per_s(images_u8, n_pairs, threads, reps=1):
    """The reference's CPU path (oracle port: same PyTorch fp32 graph as the reference modules) on
    `n_pairs` pairs of the workload with all host threads.  Test infrastructure used as the measured
    baseline only -- never on the product path."""
    import oracle
    from oracle import lightglue as olg
    from oracle import superpoint as osp
    torch.set_num_threads(threads)
    ws, wl = oracle.load_weights("superpoint_v1.pt"), oracle.load_weights("superpoint_lightglue.pt")
    lg_conf = {"depth_confidence": 0.95, "width_confidence": 0.99, "filter_threshold": 0.2, "pruning_min_kpts": -1}
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        for p in range(n_pairs):
            img = torch.from_numpy(images_u8[2 * p:2 * p + 2].astype(np.float64) / 255.0).float()[:, None]
            f0 = osp.forward(ws, img[:1], SP_CONF)
            f1 = osp.forward(ws, img[1:], SP_CONF)
            olg.forward(wl, f0["keypoints"][0][None], f0["descriptors"][0].t().contiguous()[None],
                        f1["keypoints"][0][None], f1["descriptors"][0].t().contiguous()[None], lg_conf)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_pairs / best


def run_reference(args, rank):
    if rank != 0:
        return
    sample = 2
    imgs = synth_pairs(sample)
    threads = best_cpu_threads(imgs)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1) if args.warmup < 2 else 1):
            cpu_reference_pairs_per_s(imgs, 1, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_reference_pairs_per_s(imgs, sample, threads)
        dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SuperPoint+LightGlue, synthetic 640x480 pairs (BASELINE configs[1])", "pairs_per_step": sample,
                   "max_keypoints": 1024},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} pairs/step x {args.steps} steps, reference CPU semantics (early stop + pruning every layer)"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def time_dominant_kernel(dev, tensor_cores=True):
    """CUDA-event timing of the dominant kernel -- SuperPoint conv1b (3x3, 64->64 @480x640, fused ReLU + 2x2 max-pool),
    32 images per launch as in the engine -- launched alone on the stream the bench uses.  Tensor-core path:
    tc_conv3x3_c64_kernel on pre-split bf16 planes; otherwise the fp32 CUDA-core kernel."""
    from imcui_b200 import _lib as L, ops
    lib = L.lib()
    nb = 32   # images per launch, as in the engine (imw_superpoint_forward runs the conv stack in passes of 32 images)
    x = torch.rand(nb, H, W, 64, device=dev)
    w = torch.randn(9, 64, 64, device=dev) * 0.05   # [tap][Cin][Cout]
    b = torch.zeros(64, device=dev)
    st = L.stream_ptr(dev)
    gflop = CONV1B_GFLOP_PER_IMAGE * nb
    if tensor_cores:
        # the engine's first kernel: conv1a (1 -> 64, CUDA cores inside the CTA) feeding conv1b (64 -> 64, tcgen05)
        img = torch.rand(nb, H, W, device=dev)
        w1a, b1a = torch.randn(9, 64, device=dev) * 0.3, torch.zeros(64, device=dev)
        wp = ops.split_bf16_planes(w.permute(0, 2, 1).contiguous())  # [3][tap][Cout][Cin]
        y = torch.empty(3, nb, H // 2, W // 2, 64, dtype=torch.bfloat16, device=dev)
        run = lambda: L.check(lib.imw_debug_conv1ab_fused(L.ptr(img), L.ptr(w1a), L.ptr(b1a), L.ptr(wp), L.ptr(b), L.ptr(y), nb, H, W, 1, st))
        name = "tc_conv3x3_c64_kernel<fused conv1a> (SuperPoint conv1a 1->64 + conv1b 64->64 @480x640, tcgen05 bf16x3 split = fp32-equivalent)"
        gflop += 2 * 9 * 64 * H * W * nb / 1e9
    else:
        y = torch.empty(nb, H // 2, W // 2, 64, device=dev)
        run = lambda: L.check(lib.imw_debug_conv3x3(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), nb, H, W, 64, 64, 1, 1, st))
        name = "conv3x3_nhwc_kernel (SuperPoint conv1b 64->64 @480x640, fp32 CUDA cores)"
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"ms": ms, "gflop": gflop, "images": nb, "name": name}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    assert torch.cuda.is_available(), "bench.py (b200 arm) needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    from imcui_b200 import _lib as L
    from imcui_b200.engine import PairEngine

    P = args.pairs
    lg_mode = 0 if args.fp32 else (2 if args.tf32 else 1)
    eng = PairEngine(dev, P, H, W, sp_conf={**SP_CONF, "tensor_cores": not (args.fp32 or args.sp_simt)},
                     lg_conf={"use_tensor_cores": lg_mode})
    images_u8 = synth_pairs(P, first_seed=rank * P)  # every rank matches its own shard of the pair stream
    eng.h_images.copy_(torch.from_numpy(images_u8))
    d_images = eng.to_float(eng.h_images.to(dev))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    counts_all = torch.zeros(world * P, dtype=torch.int32, device=dev) if dist is not None else None

    def gather_counts(lg):
        # the path's only exchange step: per-pair match counts to every rank
        mc = (lg["matches"][0::2] > -1).sum(1).to(torch.int32)
        if dist is not None:
            dist.all_gather_into_tensor(counts_all, mc)
            return counts_all
        return mc

    # ---- device-resident throughput (inputs already in HBM) -----------------------------------------
    for _ in range(max(args.warmup, 3)):
        sp, lg = eng.match_device(d_images)
        gather_counts(lg)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = L.lib().imw_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sp, lg = eng.match_device(d_images)
        mc = gather_counts(lg)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.lib().imw_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * P * args.steps / (ms_max / 1e3)

    # ---- end to end through the public host API (pinned host uint8 in, host results out) --------
    for _ in range(2):
        eng.match_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.match_host()
        if dist is not None:
            gather_counts(eng.lg_out)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * P * args.steps / float(t.item())

    stop = lg["stop"].float()
    n_kpts = sp["counts"][0].float()
    mean_stop, mean_kpts = float(stop.mean()), float(n_kpts.mean())
    mean_matches = float(mc.float().mean())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    roof = None
    k = time_dominant_kernel(dev, tensor_cores=not (args.fp32 or args.sp_simt))
    if k:
        ach = k["gflop"] / k["ms"]  # GFLOP/ms == TFLOP/s (algorithmic fp32 FLOPs: 2*9*Cin*Cout per output pixel)
        traffic = None
        try:  # per-launch DRAM bytes of this kernel from the committed ncu --set full capture
            traffic = json.loads((ROOT / "profiles" / "r1_conv1b_ncu.json").read_text()).get("dram_bytes_per_launch")
        except Exception:
            pass
        roof = {"kernel": k["name"], "bound": "tensor",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                "note": "split precision: six bf16 partial products per fp32-equivalent product (issued as four MMAs over N-concatenated weight planes): tensor-pipe FLOP/s = 6 x achieved",
                "peak_source": "measured bf16 burst (MEASURED_PEAKS.json)" if peaks else "fallback 1.59 PFLOP/s",
                "launch_ms": k["ms"], "algorithmic_gflop_per_launch": k["gflop"]}

    cpu = None
    if not args.no_cpu_baseline:
        threads = best_cpu_threads(images_u8)
        sample = 8
        v = cpu_reference_pairs_per_s(images_u8, sample, threads)
        cpu = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} pairs of the same synthetic stream, SuperPoint x2 + LightGlue per pair, torch CPU fp32"}

    pair_gflop = 2 * SP_GFLOP_PER_IMAGE + lg_gflop(mean_kpts, mean_stop)
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.fp32 else ("f32-equivalent: bf16x3 split tcgen05 convs (SuperPoint), " +
                                          ("single-TF32" if args.tf32 else "3xTF32 split") + " tcgen05 linears / attention / assignment (LightGlue), f32 detector post-processing"),
        "data": "synthetic",
        "config": {"workload": "SuperPoint+LightGlue, batch=64 synthetic 640x480 pairs per GPU (BASELINE configs[1])",
                   "pairs_per_gpu": P, "max_keypoints": 1024, "weights": "superpoint_v1 + GIM SP-LightGlue (real)",
                   "lightglue": "depth_confidence 0.95, width_confidence 0.99, CUDA pruning threshold 1536 (reference CUDA semantics)",
                   "mean_keypoints": mean_kpts, "mean_stop_layer": mean_stop, "mean_matches": mean_matches,
                   "l2": "per-step working set (> 1 GB of activations per 32-image conv pass) >> 126 MB L2; no explicit flush",
                   "parallelism": f"pair shard x{world}, NCCL all_gather of match counts" if world > 1 else "single GPU"},
        "achieved_tflops": value * pair_gflop / 1e3,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": eng.h2d_bytes, "d2h_bytes_per_step": eng.d2h_bytes},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
