"""Compile csrc/vats_attn.cu into csrc/libvats_attn.so for sm_100a (in-tree, so the .so travels with the snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["vats_attn.cu"]
HEADERS = ["ptx.cuh", "mask.cuh", "decode.cuh", "prefill_simt.cuh", "prefill_tc.cuh", "prefill_short.cuh", "prefill_mid.cuh", "decode_mma.cuh", "decode_prepare.cuh", "repack.cuh", "backward.cuh",
           os.path.join("..", "..", "include", "vats_attn.h")]
# VATS_BUILD_OUT: build a variant (e.g. VATS_ENABLE_TRACE=1) next to the product library instead of over it
OUT = os.environ.get("VATS_BUILD_OUT") or os.path.join(CSRC, "libvats_attn.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--use_fast_math", "-Xptxas", "-v", *([f"-DVATS_MBAR_TIMEOUT_CYCLES={os.environ['VATS_MBAR_TIMEOUT_CYCLES']}"] if os.environ.get("VATS_MBAR_TIMEOUT_CYCLES") else []),
    *(["-DVATS_ENABLE_TRACE"] if os.environ.get("VATS_ENABLE_TRACE") else []),
    *(["-DVATS_MBAR_DEBUG"] if os.environ.get("VATS_MBAR_DEBUG") else []),
    *(os.environ["VATS_EXTRA_NVCC_FLAGS"].split() if os.environ.get("VATS_EXTRA_NVCC_FLAGS") else []),   # A/B builds
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    cmd = [nvcc, *NVCC_FLAGS, "-o", OUT + ".tmp", *[os.path.join(CSRC, s) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(CSRC, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed with exit code {r.returncode} (see {log})")
    os.replace(OUT + ".tmp", OUT)
    if verbose:
        print(f"built {OUT}")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
