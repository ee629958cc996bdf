"""CPU: the reference arm of bench.py prints one JSON line with the contract's keys (the B200 arm needs
a GPU and is exercised by the driver itself)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libferox_ref_70000.so")):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bodies", "1500",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "body-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("body-steps/sec") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_ncu_captures_feed_roofline_traffic():
    """bench.py takes `roofline.traffic` (DRAM bytes of one solver launch) from the ncu summaries committed under
    profiles/: both captures must parse (kernel names with template arguments carry no comma there) and be plausible."""
    sys.path.insert(0, ROOT)
    import bench
    small, src = bench.ncu_traffic("r02_config2_top_kernels_ncu.csv", "k_solve_colored")
    large, _ = bench.ncu_traffic("r02_config3_top_kernels_ncu.csv", "k_solve_colored")
    assert src and "r02_config2_top_kernels_ncu.csv" in src
    assert 5e6 < small < 5e8, small            # 65k bodies: tens of MB, the working set lives in L2
    assert 5e8 < large < 2e10, large           # 1 M bodies: GBs, the records stream from DRAM in every sweep
    for name in ("r02_config2_top_kernels_ncu.csv", "r02_config3_top_kernels_ncu.csv"):
        rows = [l.rstrip("\n").split(",") for l in open(os.path.join(ROOT, "profiles", name))]
        assert len({len(r) for r in rows}) == 1, "ragged rows in %s: a kernel name contains a comma" % name


@pytest.mark.gpu
def test_b200_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--bodies", "4096", "--steps", "5", "--warmup", "3",
                          "--settle", "60", "--skip-sub"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert "impl" not in d and d["n_gpus"] == 1 and d["steps"] == 5 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["gpu_launches"] >= 5 * 10 and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"] * 1.05
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] == 1 and c["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
