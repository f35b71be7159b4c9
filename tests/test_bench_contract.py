"""bench.py's reference arm (the CPU leg of the measurement contract) on a tiny sample: runs here without a GPU and must
print ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_dump_fast")):
        pytest.skip("oracle/_ref not built (needs the reference sources)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                        "--ref-n", "8"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hex8 element-steps/sec fp64" and d["unit"] == "element-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["n_gpus"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_multi_gpu_roofline_object():
    """The `roofline` object of the N > 1 lines (bench.step_roofline, called by femtech_b200/dist_bench.py): the step of one
    GPU against the HBM roof, from the same byte model as the N = 1 line."""
    sys.path.insert(0, ROOT)
    import bench
    rho = (101.0 / 100.0) ** 3
    r = bench.step_roofline(8 * 4.46e9, 8, 1, True, rho)
    be, bn = bench.algorithmic_bytes(100, 1, True)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["traffic"] is None
    assert r["algorithmic_bytes_per_element"] == pytest.approx(be + bn)
    assert r["achieved"] == pytest.approx((be + bn) * 4.46e9 / 1e9)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"]) and 0 < r["frac"] < 1
    assert r["step_frac_of_hbm_roof_survey_bytes"] == pytest.approx(717.0 * 4.46 / r["peak"])
    json.dumps(r)
