"""Builds flock_b200/libflockgpu.so (CUDA kernels + C ABI + C++ host plan layer) for sm_100a, in-tree.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the gpurun snapshot.
Objects go to build/ (git-ignored); only sources whose mtime is newer than their object recompile.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "flock_b200" / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = ROOT / "flock_b200" / "libflockgpu.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas,-Wno-comment",
    "-DFLOCKGPU_BUILD",
]


CXX = os.environ.get("CXX", "g++")
CXX_FLAGS = ["-std=c++17", "-O3", "-fPIC", "-Wall", "-DFLOCKGPU_BUILD"]   # host/*.cpp: plain C++ (CPU intrinsics), no CUDA headers


def _sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cc")) + list((CSRC / "host").glob("*.cc")) + list((CSRC / "host").glob("*.cpp")))


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((CSRC / "host").glob("*.h")) + [ROOT / "include" / "flockgpu.h"]
    return max(h.stat().st_mtime for h in hs)


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    hdr = _headers_mtime()
    if obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr):
        return obj
    cmd = [NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if src.suffix == ".cpp":
        cmd = [CXX, *CXX_FLAGS, "-c", str(src), "-o", str(obj)]
    if src.suffix == ".cc":
        cmd.insert(1, "-x")
        cmd.insert(2, "cu")   # the host layer includes headers with __host__ __device__ helpers
    if verbose and src.suffix != ".cpp":
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr:
        sys.stderr.write(r.stderr)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if LIB.exists() and all(LIB.stat().st_mtime > o.stat().st_mtime for o in objs):
        return LIB
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-Xlinker", "--no-undefined", "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
