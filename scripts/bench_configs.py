"""Timings of the other BASELINE.json configs on one B200 (parity-test cases; bench.py carries the headline).

Writes one JSON object per config to stdout (and gpurun_out/configs.jsonl when that directory exists).
All timings: CUDA events on the current stream, 3 warm-ups, median of the timed repeats.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import pyaudioanalysis_b200 as pkg
from pyaudioanalysis_b200.batch import mid_pool_batch, mid_ratios, clip_stats
from oracle import st_oracle as O

pkg.ShortTermFeatures.PRINT_SPECTROGRAM_SHAPE = False
PEAK = 6487.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def noise(n_clips, n, seed=0):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return torch.randint(-12000, 12000, (n_clips, n), generator=g, device="cuda", dtype=torch.int16)


def emit(d):
    print(json.dumps(d), flush=True)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") as f:
            f.write(json.dumps(d) + "\n")


def main():
    # ---- config 1: doremi.wav through the NumPy drop-in (host API, includes malloc + copies + sync)
    g = np.load(os.path.join(ROOT, "tests", "golden", "doremi.npz"))
    x = g["x"]
    pkg.ShortTermFeatures.feature_extraction(x, 16000, 800, 400)
    t0 = time.perf_counter()
    for _ in range(20):
        F, _ = pkg.ShortTermFeatures.feature_extraction(x, 16000, 800, 400)
    dt = (time.perf_counter() - t0) / 20
    emit({"config": "1: doremi.wav 50/25 ms via ShortTermFeatures.feature_extraction (NumPy in/out, wall clock)",
          "ms": dt * 1e3, "frames": int(F.shape[1]), "frames_per_s": F.shape[1] / dt,
          "max_abs_err_vs_reference_golden": float(np.abs(F - g["st"]).max())})

    # ---- config 3: 44.1 kHz 60 s clips, 20/10 ms: spectrogram + chromagram + feature_extraction (mfcc rows)
    B3, N3, w3, s3 = 16, 2646000, 882, 441
    c3 = noise(B3, N3, 3)
    T3 = (N3 - w3) // s3 + 1
    ms_fe = timed(lambda: pkg.feature_extraction_batch(c3, 44100, w3, s3), reps=5)
    ms_sp = timed(lambda: pkg.spectrogram_batch(c3, 44100, w3, s3), reps=5)
    ms_ch = timed(lambda: pkg.chromagram_batch(c3, 44100, w3, s3), reps=5)
    alg = B3 * (2 * N3 + 4 * (T3 * 441 + (T3 - 1) * 12 + 13 * T3))
    emit({"config": "3: %d x 60 s @44.1 kHz, 20/10 ms" % B3, "frames_per_clip": T3,
          "feature_extraction_ms": ms_fe, "spectrogram_ms": ms_sp, "chromagram_ms": ms_ch,
          "feature_extraction_frames_per_s": B3 * T3 / (ms_fe * 1e-3), "spectrogram_rows_per_s": B3 * T3 / (ms_sp * 1e-3),
          "chromagram_rows_per_s": B3 * T3 / (ms_ch * 1e-3),
          "spectrogram_GBps": B3 * (2 * N3 + 4 * T3 * 441) / (ms_sp * 1e-3) / 1e9,
          "combined_algorithmic_GBps": alg / ((ms_sp + ms_ch + ms_fe) * 1e-3) / 1e9, "hbm_peak_GBps": PEAK})
    one = c3[0].cpu().numpy()[:200000]
    ref = O.spectrogram(one, 44100, w3, s3)[0]
    got = pkg.spectrogram_batch(torch.from_numpy(one).cuda()[None], 44100, w3, s3)[0].cpu().numpy()
    emit({"config": "3 parity spot check", "spectrogram_max_abs_err": float(np.abs(got - ref).max()), "ref_max": float(ref.max())})
    del c3

    # ---- config 4: mid_feature_extraction over a 1 h recording, mt 1.0/1.0 s, st 50/25 ms
    N4 = 57600000
    c4 = noise(1, N4, 4)
    ratio, stepr = mid_ratios(16000, 16000, 800, 400)

    def mid():
        st = pkg.feature_extraction_batch(c4, 16000, 800, 400)
        return mid_pool_batch(st, ratio, stepr), st
    ms4 = timed(mid, reps=5)
    m, st = mid()
    T4 = st.shape[2]
    emit({"config": "4: mid_feature_extraction, 1 h @16 kHz, mt 1.0/1.0 s, st 50/25 ms", "ms": ms4, "st_frames": int(T4),
          "mid_windows": int(m.shape[2]), "st_frames_per_s": T4 / (ms4 * 1e-3), "x_realtime": 3600.0 / (ms4 * 1e-3),
          "algorithmic_GBps": (2 * N4 + 4 * 68 * T4 + 4 * 136 * m.shape[2]) / (ms4 * 1e-3) / 1e9, "hbm_peak_GBps": PEAK})
    seg = c4[0, :480000].cpu().numpy()
    rm, rs, _ = O.mid_feature_extraction(seg, 16000, 16000, 16000, 800, 400)
    gm, gs, _ = pkg.MidTermFeatures.mid_feature_extraction(seg, 16000, 16000, 16000, 800, 400)
    emit({"config": "4 parity spot check (30 s)", "mid_max_abs_err": float(np.abs(gm - rm).max()),
          "mid_max_rel_err_over_1e-3": float((np.abs(gm - rm) / np.maximum(np.abs(rm), 1e-3)).max())})

    # ---- generic (any-window) kernel: config-2 shape with the specialised kernel disabled, and a 25/10 ms window
    from pyaudioanalysis_b200._lib import Plan
    cg = noise(1000, 160000, 5)
    pg = Plan(16000, 800, 400)
    pg.force_generic(True)
    msg = timed(lambda: pkg.feature_extraction_batch(cg, 16000, 800, 400, plan=pg), reps=5)
    ms25 = timed(lambda: pkg.feature_extraction_batch(cg, 16000, 400, 160), reps=5)            # 20x10 register-tiled kernel
    ms40 = timed(lambda: pkg.feature_extraction_batch(cg, 16000, 640, 320), reps=5)            # 20x16 register-tiled kernel
    ms64 = timed(lambda: pkg.feature_extraction_batch(cg, 16000, 1024, 512), reps=5)           # no specialisation: generic
    emit({"config": "other windows: 1000 x 10 s @16 kHz", "ms_800_400_forced_generic": msg,
          "frames_per_s_800_400_forced_generic": 399000 / (msg * 1e-3), "ms_400_160_fast_20x10": ms25,
          "frames_per_s_400_160": 1000 * ((160000 - 400) // 160 + 1) / (ms25 * 1e-3), "ms_640_320_fast_20x16": ms40,
          "frames_per_s_640_320": 1000 * ((160000 - 640) // 320 + 1) / (ms40 * 1e-3), "ms_1024_512_generic": ms64,
          "frames_per_s_1024_512": 1000 * ((160000 - 1024) // 512 + 1) / (ms64 * 1e-3)})
    del cg

    # ---- kernel 0 alone on config 2 (HBM-bound)
    c2 = noise(1000, 160000, 2)
    ms0 = timed(lambda: clip_stats(c2), reps=20)
    emit({"config": "kernel 0 (clip statistics) on 1000 x 10 s", "ms": ms0, "GBps": c2.numel() * 2 / (ms0 * 1e-3) / 1e9, "hbm_peak_GBps": PEAK})


if __name__ == "__main__":
    main()
