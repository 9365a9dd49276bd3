"""CPU: the reference arm of bench.py (`--impl reference`: the staged unmodified reference -- or the oracle port when
oracle/_ref is absent -- on the host cores) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "frames/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("audio frames/sec short-term feature_extraction")
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["vs_baseline"] is None and j["value"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and "clips" in cb["sample"]
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "pyAudioAnalysis")):
        assert cb["kind"] == "reference"          # the staged unmodified reference is what gets timed when present
    assert cb["threads_per_process"] == 1 and cb["cpu"]
    import bench
    assert j["config"] == bench.CONFIG           # both arms print the identical config dict
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]
