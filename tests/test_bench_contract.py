"""The committed bench lines (written by bench.py on the B200 box) carry every key the measurement contract names."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_single_gpu_line_has_the_contract_keys():
    d = _line("r01_bench_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"].startswith("SDS iters/sec") and d["unit"] == "it/s" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and "workload" in d["config"]
    assert abs(d["value"] - 1000.0 / d["ms_per_step"]) < 1e-6 * d["value"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"] and k in d["roofline_kernel"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert d["roofline_kernel"]["traffic"] > 0 and 0 < d["roofline_kernel"]["frac"] <= 1
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]          # a real end-to-end leg
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] > 500
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_multi_gpu_lines():
    d8, d2 = _line("r01_bench_v13_8gpu.json"), _line("r01_bench_v12_2gpu_balanced.json")
    assert d8["n_gpus"] == 8 and d2["n_gpus"] == 2 and d8["scaling"] == "strong" == d2["scaling"]
    assert d8["value"] > d2["value"] > _line("r01_bench_v13.json")["value"]
