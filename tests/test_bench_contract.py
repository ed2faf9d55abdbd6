"""CPU: the reference arm of bench.py prints one JSON line with the keys the driver reads (the GPU arm needs a
B200; its line is checked by hand against the same list in profiles/)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_line(built):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-reads", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert COMMON <= set(line), COMMON - set(line)
    assert line["metric"] == "batched_reads_per_s" and line["unit"] == "reads/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("1024 independent encrypted-address reads")


def test_committed_b200_line_has_the_contract_keys():
    """the last GPU-arm line committed under profiles/ (written by bench.py on the B200 box)"""
    files = sorted((ROOT / "profiles").glob("r2*_bench1.json"))
    assert files, "no committed bench line"
    line = json.loads(files[-1].read_text().strip().splitlines()[-1])
    assert COMMON | {"roofline", "clocks"} <= set(line)
    rf = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf)
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["scaling"] == "strong" and line["e2e"]["value"] != line["value"]
    assert {"value", "cores", "kind", "sample", "one_thread_read_s"} <= set(line["cpu_baseline"]) if line["cpu_baseline"] else True
    assert line["gpu_launches"] > 0 and not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_clock_sampler_window_selection():
    """bench.ClockSampler keeps the samples of the timed window and falls back to the whole span when the window is
    shorter than the sampling period (8-GPU runs: 0.2 s timed regions)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    row = lambda mhz, cap: [str(mhz), "1965", "700.0", "Not Active", "Not Active", "Not Active", cap]
    s.rows = [(10.0, row(1500, "Not Active")), (11.0, row(1965, "Active")), (12.0, row(1965, "Not Active")), (13.0, row(300, "Not Active"))]
    s.t0, s.t1 = 10.9, 12.1
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and out["window"] == "timed region"
    s = bench.ClockSampler(0)
    s.rows = [(10.0, row(1965, "Not Active")), (13.0, row(1965, "Not Active"))]
    s.t0, s.t1 = 11.0, 11.2
    out = s.stop()
    assert out["samples"] == 2 and out["window"].startswith("warm-up")
