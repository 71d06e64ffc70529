"""In-tree build of libb200clip.so (nvcc, sm_100a only) and of the C oracle.

Plain nvcc invocations, one object per .cu compiled in parallel, linked into
``deepcoro_clip_b200/libb200clip.so``. The .so is git-ignored but travels to the GPU box with the
gpurun snapshot. No JIT cache, no torch cpp_extension: the library has a pure C ABI (include/b200clip.h).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libb200clip.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "b200clip.h"]
    sources = sorted(CSRC.glob("*.cu"))
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_oracle(force: bool = False) -> None:
    """Prepares the checker side (never linked into the product library): the oracle itself is numpy / torch (the reference is
    Python: nothing to compile), the only artefact is oracle/_ref."""
    # oracle/_ref: the unmodified reference files of the path, copied where /root/reference exists (build container) so
    # that bench.py's reference arm / cpu_baseline legs can time the real classes on the GPU box (oracle/make_ref.py)
    rec = ROOT / "oracle" / "make_ref.py"
    if rec.exists():
        subprocess.run([sys.executable, str(rec)], check=True, capture_output=True)


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle()
    print(LIB)
