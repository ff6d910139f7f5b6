"""bench.py contract.  GPU: run `bench.py` itself on a small batch and check the printed JSON line key by key.
CPU: the host-side pieces of the bench (round-robin deal of the stream, roofline arithmetic from a launch-site profile)."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                 "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline")


def _run_bench(*extra):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "2", "--warmup", "3", *extra],
                         capture_output=True, text=True, timeout=1500, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]          # exactly ONE JSON line
    return json.loads(lines[0])


@pytest.mark.gpu
def test_bench_line_contract_config2():
    d = _run_bench("--pairs", "8", "--f1-pairs", "2", "--no-cpu-baseline")
    for k in CONTRACT_KEYS:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"] and d["data"] == "synthetic"
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and 0 < r["share_of_step"] < 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and d["gpu_launches"] > 0
    assert abs(d["value"] - 8 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6     # pairs / s from the timed steps
    assert d["match_f1"] >= 0.999 and d["parity"]["kpts_set_equal"] == 1.0 and d["parity"]["stop_equal"] == 1.0
    assert d["step_profile"]["launches"] > 50


@pytest.mark.gpu
@pytest.mark.parametrize("config,pairs", [(5, 4), (4, 4)])
def test_bench_line_contract_other_configs(config, pairs):
    d = _run_bench("--config", str(config), "--pairs", str(pairs), "--no-cpu-baseline")
    for k in CONTRACT_KEYS:
        assert k in d, k
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0 and d["roofline"].get("frac", 0) > 0


def test_round_robin_deal_covers_the_stream():
    sys.path.insert(0, str(ROOT))
    import bench
    for world in (1, 2, 4, 8):
        for rank in range(world):
            seen = [(g * world + rank) % bench.NB for g in range(bench.NB)]
            assert sorted(seen) == list(range(bench.NB)), (world, rank, seen)     # every rank meets every batch once per cycle


def test_roofline_arithmetic_from_a_site_profile():
    sys.path.insert(0, str(ROOT))
    import bench

    class Prof:
        sites = [("int tc_conv1ab_fused(...):597", 4, 16.0), ("int launch_tc_attn(...):266", 10, 4.0)]
        total_ms = 40.0

        def find(self, *needles):
            return [t for t in self.sites if all(n in t[0] for n in needles)]

    r = bench.roofline_from_sites(Prof(), ["tc_conv1ab_fused"], 2944.0, "TFLOP/s", "tensor", "k")     # 2944 GFLOP in 16 ms
    assert abs(r["achieved"] - 184.0) < 1e-9 and abs(r["frac"] - 184.0 / r["peak"]) < 1e-12
    assert r["launches_per_step"] == 4 and abs(r["share_of_step"] - 0.4) < 1e-12 and abs(r["launch_ms"] - 4.0) < 1e-12
    h = bench.roofline_from_sites(Prof(), ["launch_tc_attn"], 8.0, "GB/s", "hbm", "k")                 # 8 GB in 4 ms = 2000 GB/s
    assert abs(h["achieved"] - 2000.0) < 1e-9
    assert "error" in bench.roofline_from_sites(Prof(), ["nope"], 1.0, "GB/s", "hbm", "k")
