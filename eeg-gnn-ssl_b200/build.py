"""Build libdcgru_b200.so (sm_100a) in-tree with nvcc.  No torch involved: the library is a plain
C-ABI shared object (include/dcgru_b200.h) that the Python side loads with ctypes."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libdcgru_b200.so")
SOURCES = ["seq_fwd.cu", "seq_bwd.cu", "dw.cu", "graph.cu", "tc_selftest.cu", "dw_tc.cu", "dw_mm.cu", "optim.cu", "seq_fwd_tc.cu", "seq_bwd_tc.cu", "tmap.cu", "bulk_dp.cu", "rnn_fwd.cu", "rnn_bwd.cu", "dw_mm16.cu", "head.cu", "fft.cu", "capi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


JITTER_LIB = os.path.join(OUT_DIR, "libdcgru_b200_jitter.so")


def build(verbose=False, force=False, jitter=False):
    """jitter=True: the stress variant libdcgru_b200_jitter.so (-DDCGRU_JITTER=3000: random sleeps in front of every mbarrier
    operation, csrc/tc_common.cuh), objects under lib/jitter/; used by tests/test_gpu_stress.py only."""
    if jitter:
        return _build(verbose, force, os.path.join(OUT_DIR, "jitter"), JITTER_LIB, ["-DDCGRU_JITTER=3000"])
    return _build(verbose, force, OUT_DIR, LIB, [])


def _build(verbose, force, obj_dir, lib_path, extra):
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "dcgru_b200.h"))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd[-4:]) + "\n" + r.stdout + r.stderr)
        return r.returncode

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            rcs = list(ex.map(run, jobs))
        if any(rcs):
            raise RuntimeError("nvcc failed")
    if jobs or force or _stale(lib_path, objs):
        cmd = [NVCC, "-shared", "-o", lib_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return lib_path


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, jitter="--jitter" in sys.argv))
