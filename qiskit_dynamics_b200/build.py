"""Builds csrc/*.cu into csrc/libqdb.so with nvcc for sm_100a (in-tree, so it ships with gpurun)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "gen.cu", "zgemm.cu", "rk4_fused.cu", "rk4_rowsplit.cu", "rk4_ozaki.cu", "zgemm_ozaki.cu", "rk4_sweep_small.cu", "rk4_sweepf.cu", "lindblad.cu", "expm.cu", "propagator.cu", "signals.cu", "measure.cu"]
OUT = os.path.join(CSRC, "libqdb.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--threads", "0",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "qdb.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("QDB_NVCC_EXTRA", "").split()  # e.g. -DQDB_OZ_TIMELINE for profiles/probe/ozaki_probe.py
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libqdb.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build_library(force=True, verbose="-v" in sys.argv))
