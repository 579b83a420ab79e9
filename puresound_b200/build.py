"""Build the sm_100a shared library in-tree (``puresound_b200/libpuresound_b200.so``).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU
box with the gpurun snapshot.  ``python -m puresound_b200.build`` rebuilds.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpuresound_b200.so")
SOURCES = ["ps_api.cu", "ps_gemm_simt.cu", "ps_gemm_tc.cu", "ps_gemm_pair.cu", "ps_gemm_wide.cu", "ps_gemm_wide_tma.cu", "ps_gemm_rows.cu", "ps_norm.cu", "ps_dwconv.cu", "ps_dwconv_tma.cu", "ps_misc.cu", "ps_gated.cu", "ps_attention.cu", "ps_sdr.cu", "ps_lstm.cu", "ps_lstm_tc.cu", "ps_stream.cu", "ps_stream_hop.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode() + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """experiments=True builds libpuresound_b200_exp.so with -DPS_EXPERIMENTS (the bottleneck switches of the GEMM / LSTM
    kernels: results are garbage) next to the release library; it is only ever loaded through PS_B200_LIB (profiling runs)."""
    global OBJ, LIB
    if experiments:
        OBJ, LIB = os.path.join(HERE, "build_exp"), os.path.join(HERE, "libpuresound_b200_exp.so")
    else:
        OBJ, LIB = os.path.join(HERE, "build"), os.path.join(HERE, "libpuresound_b200.so")
    flags = NVCC_FLAGS + (["-DPS_EXPERIMENTS"] if experiments else [])
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest() + ("exp" if experiments else "")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".ptxas.log", "w") as fh:
            fh.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv))
