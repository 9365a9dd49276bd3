"""Small driver for compute-sanitizer: every kernel family once on tiny inputs (features fast/generic, rows, mid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pyaudioanalysis_b200 as pkg
from oracle import st_oracle as O

pkg.ShortTermFeatures.PRINT_SPECTROGRAM_SHAPE = False
clips = torch.from_numpy(np.stack([O.synth_clip(i, 24000, 16000) for i in range(6)])).cuda()
out = pkg.feature_extraction_batch(clips, 16000, 800, 400)                       # fast kernel, RUNS + TMA
lens = torch.tensor([24000, 800, 12345, 23999, 20000, 4000], dtype=torch.int64, device="cuda")
pkg.feature_extraction_batch(clips, 16000, 800, 400, lengths=lens)               # ragged
pkg.feature_extraction_batch(clips[:, 3:23003].contiguous(), 16000, 800, 400)    # same kernel, other lengths
pkg.feature_extraction_batch(clips, 16000, 800, 200)                              # 75 % overlap: no Zs aliasing
pkg.feature_extraction_batch(clips, 16000, 800, 800)                              # no overlap
c44 = torch.from_numpy(np.stack([O.synth_clip(9 + i, 30000, 44100) for i in range(3)])).cuda()
pkg.feature_extraction_batch(c44, 44100, 882, 441)                                # R = 21, odd hop
pkg.feature_extraction_batch(clips, 16000, 640, 160)                              # 20x16 (Cooley-Tukey 16-point codelet)
pkg.feature_extraction_batch(clips, 16000, 320, 160)                              # 16x10
pkg.feature_extraction_batch(clips, 16000, 1024, 512)                             # generic kernel
pkg.feature_extraction_batch(clips, 16000, 400, 160)                              # 20x10 rectangular, run staging
pkg.feature_extraction_batch(clips, 16000, 480, 240)                              # 20x12
pkg.feature_extraction_batch(clips, 16000, 600, 150)                              # 20x15, no runs
pkg.spectrogram_batch(clips, 16000, 400, 160)
pkg.spectrogram_batch(clips, 16000, 800, 400)
pkg.chromagram_batch(clips, 16000, 800, 400)
pkg.spectrogram_batch(c44, 44100, 882, 441)
pkg.mid_feature_extraction_batch(clips, 16000, 8000, 4000, 800, 400)
x = clips[0].cpu().numpy()
pkg.ShortTermFeatures.chromagram(x[:16300], 16000, 800, 400)                      # clipped last frame (generic launch)
torch.cuda.synchronize()
print("sanitize driver done", tuple(out.shape))
