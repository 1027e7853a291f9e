"""Builds libsegofa_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ["runtime.cu", "gemm.cu", "gemm_ts.cu", "attention.cu", "attention_bwd.cu", "rowops.cu", "train.cu", "segmask.cu", "preprocess.cu"]
LIB = os.path.join(PKG, "libsegofa_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(HERE)) + ["../../include/segofa_b200.h"]:
        p = os.path.join(HERE, f)
        if os.path.isfile(p) and p.endswith((".cu", ".cuh", ".h", ".py")):
            h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = LIB + ".stamp"
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objs = []
    logs = []
    procs = []
    for src in SOURCES:  # one nvcc per translation unit, all at once (gemm.cu alone takes minutes)
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = None
    for src, cmd, pr in procs:
        out, _ = pr.communicate()
        logs.append(f"$ {' '.join(cmd)}\n{out}")
        if pr.returncode != 0 and failed is None:
            failed = src
            sys.stderr.write(logs[-1])
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(HERE, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
