"""The bench line committed under profiles/ carries every key of the bench contract (guards bench.py edits)."""
import json

from conftest import ROOT


def _line(name):
    return json.loads((ROOT / "profiles" / name).read_text().strip().splitlines()[-1])


def test_bench_line_has_the_contract_keys():
    d = _line("r1_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]          # the end-to-end number includes the copies
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert set(c) >= {"value", "unit", "cores", "kind", "sample"} and c["kind"] in ("port", "reference")
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and d["gpu_launches"] > 0
    assert abs(d["value"] - 64 * d["n_gpus"] / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6   # pairs / s from the timed steps


def test_two_gpu_line_scales():
    d1, d2 = _line("r1_bench_n1.json"), _line("r1_bench_n2.json")
    assert d2["n_gpus"] == 2 and d2["metric"] == d1["metric"] and d2["value"] > 1.8 * d1["value"]
