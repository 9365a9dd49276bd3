"""Tiny driver for ncu: a few launches of the hot path on the bench workload (no timing claims)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pyaudioanalysis_b200 as pkg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
clips = bench.synth_device_batch(torch, n, 1234, torch.device("cuda", 0))
out = None
for _ in range(reps):
    norm = pkg.clip_stats(clips)
    out = pkg.feature_extraction_batch(clips, bench.FS, bench.WINDOW, bench.STEP, norm=norm, out=out)
torch.cuda.synchronize()
print("done", tuple(out.shape))
