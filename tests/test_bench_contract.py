"""bench.py contract (CPU part): the reference arm runs without a GPU, prints ONE JSON line with the keys the driver
reads, and never touches /root/reference; the native arm refuses to run without the CUDA library path (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GDOF*steps/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


def test_bench_sources_do_not_read_the_reference_tree():
    for f in ("bench.py", "__graft_entry__.py"):
        assert "/root/reference" not in open(os.path.join(ROOT, f)).read(), f


def test_headline_constants_and_committed_line():
    """The headline mesh is BASELINE.md's ~100k-element config, the measured-traffic table is keyed by its size, and the
    committed final bench line carries every key of the contract (roofline with measured traffic, cpu_baseline, e2e with host
    copies, clocks, launch count, the same config keys as the reference arm)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.LAYERS_HEADLINE * 1996 == 99800
    assert list(bench.NCU_TRAFFIC_AXHELM8A) == [99800 * 512]
    line = [l for l in open(os.path.join(ROOT, "profiles", "r02_bench_headline_final.json")) if l.startswith("{")][-1]
    d = json.loads(line)
    ref = json.loads([l for l in open(os.path.join(ROOT, "profiles", "r02_bench_reference_arm.json")) if l.startswith("{")][-1])
    assert d["config"]["elements"] == 99800 and d["n_gpus"] == 1 and d["scaling"] == "strong" and d["dtype"] == "f64"
    for k in ("workload", "elements", "lx1", "lxd", "layers_total", "timestepper", "residualProj", "tolerances"):
        assert d["config"][k] == ref["config"][k], k                       # same workload in both arms
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1       # measured DRAM bytes ~ algorithmic bytes
    assert r["us_per_launch"] * r["launches_per_step"] < d["ms_per_step"] * 1e3   # dominant-kernel time < step time
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] <= d["value"] * 1.02
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["cpu_baseline"]["kind"] == "port" and d["cylinder_re50"]["time_to_leading_eigs_s"] > 0
    # whole-step sanity: bytes the kernels of one step must move at least / step time stays below the HBM peak
    assert r["algorithmic_bytes_per_launch"] * r["launches_per_step"] / (d["ms_per_step"] * 1e-3) / 1e9 < r["peak"]
