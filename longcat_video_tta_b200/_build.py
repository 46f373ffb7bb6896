"""In-tree build of libb200tta.so (nvcc, sm_100a only).  No JIT cache: the built library sits next to the
sources so that it travels with a snapshot of the repository."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "build"
LIB = PKG / "libb200tta.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (Path(c).exists() or c == "nvcc"):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "b200tta.h"]
    return max(h.stat().st_mtime for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
    # objects of a developer build (B200TTA_ATTN_DEBUG: timing counters in the attention kernels) live in their own
    # directory and the library records which flavour it was linked from, so a stale debug object can never end up in
    # a release library (or the other way round)
    debug = bool(os.environ.get("B200TTA_ATTN_DEBUG"))
    obj_dir = CSRC / ("build_debug" if debug else "build")
    obj_dir.mkdir(parents=True, exist_ok=True)
    stamp = OBJ / "linked_flavour"
    flavour = "debug" if debug else "release"
    relink = not stamp.exists() or stamp.read_text() != flavour
    nvcc = _nvcc()
    hdr_m = _deps_mtime()
    jobs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if debug:
            cmd.insert(1, "-DB200TTA_ATTN_DEBUG=1")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(compile_one, jobs):
                if verbose and out:
                    print(out, file=sys.stderr)
    objs = [obj_dir / (s.stem + ".o") for s in sources()]
    if jobs or relink or not LIB.exists():
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        OBJ.mkdir(parents=True, exist_ok=True)
        stamp.write_text(flavour)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
