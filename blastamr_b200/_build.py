"""In-tree builds: the CUDA C-ABI library (nvcc, sm_100a), the host mesh engine (gcc)
and -- for tests / the CPU baseline only -- the oracle (gcc).

Everything is compiled to a shared object next to its sources so that the built files
travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "blastamr_b200" / "csrc"
LIB_CUDA = ROOT / "blastamr_b200" / "libamr_b200.so"
LIB_MESH = ROOT / "blastamr_b200" / "libamrmesh.so"
LIB_HOST = ROOT / "blastamr_b200" / "libamrhost.so"
ORACLE_DIR = ROOT / "oracle"
LIB_ORACLE = ORACLE_DIR / "liboracle.so"

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


class _Lock:
    """one builder at a time per target (ranks of one torchrun launch call the builders concurrently); the output is written
    next to the target and renamed into place, so a loader never sees a half-written library"""

    def __init__(self, target: Path):
        self.path = target.with_suffix(target.suffix + ".lock")
        self.fh = None

    def __enter__(self):
        import fcntl
        self.fh = open(self.path, "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()
        return False


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(map(str, cmd)) + "\n" + r.stdout + "\n")
        raise RuntimeError(f"build failed: {cmd[0]} (exit {r.returncode})")
    return r.stdout


def _nccl_paths():
    """Prefer the NCCL that torch bundles (the one the process has loaded anyway)."""
    try:
        import nvidia.nccl as _n  # type: ignore

        base = Path(list(_n.__path__)[0])
        inc, lib = base / "include", base / "lib"
        if (inc / "nccl.h").exists() and (lib / "libnccl.so.2").exists():
            return str(inc), str(lib), "libnccl.so.2"
    except Exception:
        pass
    return "/usr/include", "/usr/lib/x86_64-linux-gnu", "libnccl.so.2"


def build_mesh(force: bool = False) -> Path:
    src = [CSRC / "meshgen.c"]
    if not force and _newer(LIB_MESH, src):
        return LIB_MESH
    with _Lock(LIB_MESH):
        if not force and _newer(LIB_MESH, src):
            return LIB_MESH
        tmp = LIB_MESH.with_suffix(f".tmp{os.getpid()}.so")
        _run(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden", "-std=gnu11",
              "-o", str(tmp)] + [str(s) for s in src] + ["-lm"])
        os.replace(tmp, LIB_MESH)
    return LIB_MESH


def build_oracle(force: bool = False) -> Path:
    src = sorted(ORACLE_DIR.glob("*.c")) + sorted(ORACLE_DIR.glob("*.h"))
    csrc = [s for s in src if s.suffix == ".c"]
    if not force and _newer(LIB_ORACLE, src):
        return LIB_ORACLE
    with _Lock(LIB_ORACLE):
        if not force and _newer(LIB_ORACLE, src):
            return LIB_ORACLE
        tmp = LIB_ORACLE.with_suffix(f".tmp{os.getpid()}.so")
        # -ffp-contract=off: no FMA contraction, x86-64 OpenFOAM builds have none either
        _run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-std=gnu11", "-pthread",
              "-o", str(tmp)] + [str(s) for s in csrc] + ["-lm"])
        os.replace(tmp, LIB_ORACLE)
    return LIB_ORACLE


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    src = sorted(CSRC.glob("*.cu"))
    hdr = sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    if not force and _newer(LIB_CUDA, src + hdr):
        return LIB_CUDA
    with _Lock(LIB_CUDA):
        if not force and _newer(LIB_CUDA, src + hdr):      # another rank built it while this one waited
            return LIB_CUDA
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        inc, libdir, libname = _nccl_paths()
        tmp = LIB_CUDA.with_suffix(f".tmp{os.getpid()}.so")
        cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", *NVCC_ARCH,
               "-fmad=false",                      # bitwise parity of the FP64 estimators with the CPU reference
               "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
               "-I", str(ROOT / "include"), "-I", inc,
               "-o", str(tmp)] + [str(s) for s in src] + [
               "-L", libdir, f"-l:{libname}", "-Xlinker", f"-rpath={libdir}", "-lcudart"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        out = _run(cmd)
        os.replace(tmp, LIB_CUDA)
        if verbose:
            print(out)
    return LIB_CUDA


def build_host(force: bool = False) -> Path:
    src = sorted((ROOT / "blastamr_b200" / "host").glob("*.cpp"))
    hdr = sorted((ROOT / "blastamr_b200" / "host").glob("*.H")) + sorted((ROOT / "include").glob("*.h"))
    if not src:
        return LIB_HOST
    if not force and _newer(LIB_HOST, src + hdr + [LIB_CUDA]):
        return LIB_HOST
    with _Lock(LIB_HOST):
        if not force and _newer(LIB_HOST, src + hdr + [LIB_CUDA]):
            return LIB_HOST
        tmp = LIB_HOST.with_suffix(f".tmp{os.getpid()}.so")
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I", str(ROOT / "include"),
              "-I", str(ROOT / "blastamr_b200" / "host"), "-o", str(tmp)] + [str(s) for s in src] +
             ["-L", str(LIB_CUDA.parent), "-l:libamr_b200.so", "-Wl,-rpath,$ORIGIN"])
        os.replace(tmp, LIB_HOST)
    return LIB_HOST


def build_all(force: bool = False):
    build_mesh(force)
    build_cuda(force)
    build_host(force)
    if ORACLE_DIR.exists():
        build_oracle(force)        # (no oracle/_ref: the reference needs OpenFOAM + wmake and cannot be compiled here)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB_MESH, LIB_CUDA, LIB_ORACLE)
