"""bench.py's reference arm (`--impl reference`: the C port of the reference's loop on the host cores) needs no GPU: it
must print ONE JSON line with the keys the driver reads, under the same metric / unit / direction as the GPU arm."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "kzg_commit_msm_points_per_s" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


def test_gpu_arm_names_the_same_metric_and_declares_its_keys():
    """static check of the GPU arm (it needs a device to run): same metric string, and every key of the contract is
    assembled in the line"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"kzg_commit_msm_points_per_s"') >= 2
    for key in ("ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches",
                "clocks", "roofline", "cpu_baseline", "h2d_bytes_per_step", "d2h_bytes_per_step", "calls_in_flight", "fused_calls"):
        assert re.search(r'"%s"\s*:' % re.escape(key), src), key
