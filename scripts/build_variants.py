"""Build experimental variants of libb200aa.so next to the default one (A/B runs on the GPU box).

    python scripts/build_variants.py            # all variants
    python scripts/build_variants.py lean       # one

Variants land in pyaudioanalysis_b200/variants/libb200aa_<name>.so (git-ignored like every .so, shipped to the
GPU box by gpurun) and are selected per process with the B200AA_LIB environment variable.  One gpurun call A/Bs
them (quick parity against the oracle + kernel timing per build):

    gpurun --timeout 300 -- 'python scripts/ab_run.py default lean lean6 mb4 hostpipe | tee gpurun_out/ab.jsonl'

and the full suite runs on a variant with `B200AA_LIB=$PWD/pyaudioanalysis_b200/variants/libb200aa_lean.so python -m
pytest tests -m gpu -q`.

The default build is never affected: every variant is a compile-time switch that is off by default (the SASS of
the default library was compared before / after the switches were added).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyaudioanalysis_b200 import build as B   # noqa: E402

VARIANTS = {
    # 4 CTAs / SM for the run-staged feature kernels: 64 registers, 56 KB shared memory (tables through L1,
    # one carried |X| row, 16-bit flip words); see B200AA_FAST_LEAN in csrc/fast_kernel.cuh
    "lean": ["-DB200AA_FAST_LEAN=1"],
    # lean + six-warp CTAs (five transform warps + one spare, dense pass on four warps, mel / chroma on two):
    # 4 CTAs / SM at 80 registers instead of 64
    "lean6": ["-DB200AA_FAST_LEAN=2"],
    # 64 registers / 4 CTAs per SM without the diet: only the small-window shapes (<= 45 KB: 320 / 400 / 480-sample
    # windows) actually reach 4 CTAs per SM with it
    "mb4": ["-DB200AA_FAST_MINBLOCKS=4"],
    # b200aa_st_features_host as a chunked three-stream pipeline below the C ABI (end-to-end through the C entry point)
    "hostpipe": ["-DB200AA_HOST_PIPELINE=1"],
    # reference points for bisecting: scalar butterflies / IEEE MUFU wrappers
    "nof32x2": ["-DB200AA_NO_F32X2"],
    "noftz": ["-DB200AA_NO_FTZ_MUFU"],
}


def build(name, verbose=False):
    out_dir = os.path.join(B.HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libb200aa_%s.so" % name)
    flags = [f for f in B.NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [B._nvcc()] + flags + VARIANTS[name] + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out] + [os.path.join(B.CSRC, s) for s in B.SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building variant " + name)
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(VARIANTS)
    for n in names:
        print(build(n, verbose="-v" in sys.argv))
