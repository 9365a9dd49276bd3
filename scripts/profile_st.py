"""Tiny driver for ncu: a few launches of the hot path on the bench workload shape (no timing claims).
Input is plain seeded noise (one generator kernel) so ncu does not spend its time on data generation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pyaudioanalysis_b200 as pkg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda")
g.manual_seed(1234)
clips = torch.randint(-12000, 12000, (n, bench.CLIP_SAMPLES), generator=g, device="cuda", dtype=torch.int16)
out = None
for _ in range(reps):
    norm = pkg.clip_stats(clips)
    out = pkg.feature_extraction_batch(clips, bench.FS, bench.WINDOW, bench.STEP, norm=norm, out=out)
torch.cuda.synchronize()
print("done", tuple(out.shape))
