"""Tiny driver for ncu: a few launches of the hot path on the bench workload shape, or `n reps fs window step n_samples
[features|spectrogram|chromagram]` (no timing claims).
Input is plain seeded noise (one generator kernel) so ncu does not spend its time on data generation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pyaudioanalysis_b200 as pkg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fs, w, s, ns = (int(a) for a in sys.argv[3:7]) if len(sys.argv) > 6 else (bench.FS, bench.WINDOW, bench.STEP, bench.CLIP_SAMPLES)
mode = sys.argv[7] if len(sys.argv) > 7 else "features"
g = torch.Generator(device="cuda")
g.manual_seed(1234)
clips = torch.randint(-12000, 12000, (n, ns), generator=g, device="cuda", dtype=torch.int16)
out = None
for _ in range(reps):
    norm = pkg.clip_stats(clips)
    if mode == "spectrogram":
        out = pkg.spectrogram_batch(clips, fs, w, s)
    elif mode == "chromagram":
        out = pkg.chromagram_batch(clips, fs, w, s)
    else:
        out = pkg.feature_extraction_batch(clips, fs, w, s, norm=norm, out=out)
torch.cuda.synchronize()
print("done", tuple(out.shape))
