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
