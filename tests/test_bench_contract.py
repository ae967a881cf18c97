"""bench.py's output contract on the arm that runs without a GPU: exactly one line on stdout, and that line is the JSON object
(the compiled reference and NCCL both print to file descriptor 1 from native code; bench.py moves all of that to stderr)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    from oracle import ref_binding
    if not ref_binding.available("hdl64_1800"):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-sweeps", "2"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout[:2000]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "scans/s" and line["value"] > 0
    assert line["steps"] == 1 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"]
