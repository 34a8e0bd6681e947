"""Builds librmdf_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python build.py [--force] [--verbose]

Per-file flags matter:
  * march_parity.cu, julia.cu, envmap.cu, rmdf_api.cu: -fmad=false  (IEEE-only arithmetic
    so device results are bit-identical to the CPU oracle; host code -ffp-contract=off)
  * march_fast.cu: --use_fast_math
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "librmdf_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math"]
UNITS = {
    "rmdf_api.cu": ["-fmad=false"],
    "march_parity.cu": ["-fmad=false"],
    "march_fast.cu": ["--use_fast_math"],
    "envmap.cu": ["-fmad=false"],
    "julia.cu": ["-fmad=false"],
    "frame_sync.cu": [],
    "peak.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    return list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((HERE.parent / "include").glob("*.h"))


def build(force: bool = False, verbose: bool = False, extra=(), out: Path = LIB, objdir: Path = OBJ) -> Path:
    """extra: additional nvcc flags (e.g. -DRMDF_MIN_CTAS=3) for tuning variants; out/objdir
    let a variant live beside the default library (RMDF_LIB selects it at load time)."""
    global OBJ, LIB
    nvcc = _nvcc()
    OBJ, LIB = Path(objdir), Path(out)
    OBJ.mkdir(exist_ok=True)
    newest_dep = max(p.stat().st_mtime for p in _deps() + [Path(__file__)])
    objs, rebuilt = [], False
    for name, flags in UNITS.items():
        src, obj = CSRC / name, OBJ / (name + ".o")
        objs.append(str(obj))
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, newest_dep):
            continue
        cmd = [nvcc] + COMMON + flags + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s" % name)
        rebuilt = True
    if rebuilt or force or not LIB.exists():
        r = subprocess.run([nvcc] + ARCH + ["-shared", "-o", str(LIB)] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--variant", default="", help="name of a tuning variant: builds librmdf_b200_<name>.so")
    ap.add_argument("--extra", default="", help="extra nvcc flags, space separated")
    a = ap.parse_args()
    if a.variant:
        print(build(True, a.verbose, a.extra.split(), HERE / ("librmdf_b200_%s.so" % a.variant), HERE / ("build_" + a.variant)))
    else:
        print(build(a.force, a.verbose, a.extra.split()))
