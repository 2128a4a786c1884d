"""bench.py's JSON contract, checked on the CPU tier through the reference arm (the GPU arm needs a device): one JSON line
with the base keys, the same `metric` / `config` as the GPU arm, the `cpu_baseline` and `e2e` objects the tier asks for."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sample-n", "512"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["sample_n"] == 512 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the reference arm runs on the GPU arm's workload description
    sys.path.insert(0, ROOT)
    import bench
    assert d["metric"] == bench.METRIC and d["config"] == bench.workload_config(32768, 1)
    assert "N=32768" in d["config"]["workload"] and "inputs_exceed_l2" in d["config"]["l2"]
    assert d["ms_per_step"] < d["ms_per_step_full_size_extrapolated"]


def test_sharded_bench_module_is_importable_without_a_gpu():
    sys.path.insert(0, ROOT)
    import bench_sharded
    assert bench_sharded.VFE_N == 10_000_000 and bench_sharded.VFE_M == 1024 and bench_sharded.SVGP_M == 2048
    assert bench_sharded.VFE_SHARDS % 8 == 0 and bench_sharded.C5_N == 131072
