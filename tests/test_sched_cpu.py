"""CPU: the work distribution of the warp-autonomous kernels (csrc/sched.cuh: static shares + steal-half) run with one
thread per "warp" -- every pair step of a launch is claimed exactly once whatever the interleaving, runs stay contiguous
(one halo per warp and per steal), degenerate launches (fewer pairs than warps, no pairs) terminate."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_steal_half_scheduler(tmp_path):
    exe = str(tmp_path / "sched_host")
    res = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-x", "c++", "-o", exe, os.path.join(ROOT, "tests", "sched_host.cpp")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # n_warps total chunk min_steal slow_every
    for args in ((64, 100000, 4, 6, 5), (200, 200000, 4, 6, 3), (37, 1000, 2, 4, 2), (128, 50, 4, 6, 0), (16, 0, 4, 6, 0),
                 (300, 300000, 1, 2, 7), (8, 1000000, 4, 6, 1), (96, 2960 * 67, 4, 6, 4)):
        out = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, (args, out.stdout, out.stderr)
        r = json.loads(out.stdout)
        assert r["missed"] == 0 and r["duplicated"] == 0, r
        # a run starts where a warp's share starts or where it stole: nothing else breaks contiguity
        assert r["run_starts"] <= min(args[0], max(args[1], 1)) + r["steals"], r
