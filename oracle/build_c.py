"""Compile oracle/dft_def.c -> oracle/_build/libdftdef.so (TEST INFRASTRUCTURE ONLY)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libdftdef.so")


def build(force=False):
    src = os.path.join(HERE, "dft_def.c")
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(src):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", SO, src, "-lm"])
    return SO


if __name__ == "__main__":
    print(build(force=True))
