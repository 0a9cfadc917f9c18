"""In-tree builds (no JIT cache): the CUDA library, the host-side scene/narrow-phase fixture library.

`nvcc` cross-compiles sm_100a without a GPU, so this runs on the CPU-only authoring box; the resulting
`.so` files travel to the GPU box with the repo snapshot (they are git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
LIB_DIR = ROOT / "lib"
CUDA_LIB = LIB_DIR / "libavian_b200.so"
HOST_LIB = LIB_DIR / "libavian_host.so"

CUDA_SOURCES = ["abi.cu", "solver_host.cu", "broadphase.cu", "aabb.cu", "narrow.cu", "contacts.cu"]
CUDA_HEADERS = ["avn_math.cuh", "solver_dev.cuh", "joints_dev.cuh", "solver_kernels.cuh", "context.hpp", "joint_schedule.hpp", "broadphase_cells.cuh", "narrow_math.hpp", "contact_rows.hpp"]
# -fmad=false: the reference (Rust) never contracts a*b+c; parity at 1e-5 on contact dynamics needs the same
# rounding.  Division and sqrt stay IEEE (nvcc defaults -prec-div=true -prec-sqrt=true).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]
HOST_SOURCES = ["host_api.cpp"]
HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall"]


def _newer(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(d.stat().st_mtime <= t for d in deps if d.exists())


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    return None


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    src = ROOT / "csrc"
    deps = [src / s for s in CUDA_SOURCES + CUDA_HEADERS] + [REPO / "include" / "avian_b200.h"]
    if not force and _newer(CUDA_LIB, deps):
        return CUDA_LIB
    nvcc = find_nvcc()
    if nvcc is None:
        if CUDA_LIB.exists():
            return CUDA_LIB  # prebuilt library shipped with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libavian_b200.so: the CUDA path cannot be built")
    LIB_DIR.mkdir(exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(CUDA_LIB), *[str(src / s) for s in CUDA_SOURCES]]
    subprocess.run(cmd, check=True, cwd=src)
    return CUDA_LIB


def build_host(force: bool = False) -> Path:
    src = ROOT / "host"
    deps = [p for p in src.glob("*.[ch]pp")] + [REPO / "include" / "avian_b200.h", ROOT / "csrc" / "narrow_math.hpp", ROOT / "csrc" / "contact_rows.hpp"]
    if not force and _newer(HOST_LIB, deps):
        return HOST_LIB
    cxx = os.environ.get("CXX") or shutil.which("g++")
    if cxx is None:
        if HOST_LIB.exists():
            return HOST_LIB
        raise RuntimeError("g++ not found and no prebuilt libavian_host.so")
    LIB_DIR.mkdir(exist_ok=True)
    subprocess.run([cxx, *HOST_FLAGS, "-o", str(HOST_LIB), *[str(src / s) for s in HOST_SOURCES]], check=True, cwd=src)
    return HOST_LIB


def build_oracle(force: bool = False) -> Path:
    """Test infrastructure only (tests/, smoke(), bench.py cpu_baseline)."""
    odir = REPO / "oracle"
    target = odir / "_build" / "liboracle.so"
    if force and target.exists():
        target.unlink()
    if shutil.which("make") and (shutil.which("g++") or os.environ.get("CXX")):
        subprocess.run(["make", "-s"], check=True, cwd=odir)
    if not target.exists():
        raise RuntimeError("oracle/_build/liboracle.so missing and cannot be built")
    return target


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)
    build_oracle(force)
