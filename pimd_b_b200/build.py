"""Build libpimdb200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python -m pimd_b_b200.build [--force] [--verbose]

No JIT cache, no torch dependency: the shared object lands next to this file so it travels with the
repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libpimdb200.so"
SOURCES = ["api.cu", "pair_forces.cu", "exchange.cu", "integrator.cu", "normal_modes.cu", "nose_hoover.cu", "ranmars.cu", "factorial.cu"]
HEADERS = [CSRC / "internal.cuh", CSRC / "device_utils.cuh", PKG.parent / "include" / "pimdb200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libpimdb200.so cannot be built")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = nvcc_path()
    OBJ.mkdir(exist_ok=True)
    jobs = []
    for src in SOURCES:
        s = CSRC / src
        o = OBJ / (s.stem + ".o")
        if force or _stale(o, [s, *HEADERS]):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out)
    objs = [str(OBJ / (Path(s).stem + ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return LIB


HOST_DIR = PKG / "host"
HOST_BIN = PKG / "pimdb_gpu"


def build_host(force: bool = False) -> Path:
    """pimdb_gpu: the C++20 host mirror of the reference's plugin surface + entry point, linked against the C ABI."""
    srcs = [HOST_DIR / "pimdb_host.cpp", HOST_DIR / "pimdb_gpu.cpp"]
    deps = srcs + [HOST_DIR / "pimdb_host.hpp", PKG.parent / "include" / "pimdb200.h", LIB]
    if force or _stale(HOST_BIN, deps):
        cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
        cuda_lib = str(Path(nvcc_path()).resolve().parent.parent / "lib64")
        cmd = [cxx, "-std=c++20", "-O2", "-o", str(HOST_BIN), *map(str, srcs), f"-L{PKG}", "-lpimdb200",
               "-Wl,-rpath,$ORIGIN", f"-L{cuda_lib}", f"-Wl,-rpath,{cuda_lib}", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return HOST_BIN


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
    print(build_host(force="--force" in sys.argv))
