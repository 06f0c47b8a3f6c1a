#!/usr/bin/env python
"""Build libgf3b200.so (sm_100a only) in-tree: gf3-audio-modem_b200/lib/libgf3b200.so.

    python gf3-audio-modem_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  One object per .cu (compiled in parallel), linked with
`nvcc -shared`.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, os.environ.get("GF3_LIB_NAME", "libgf3b200.so"))   # experiments: alternate name + flags
SOURCES = ["gf3_lib.cu", "gf3_rx.cu", "gf3_rx_staged_u8.cu", "gf3_rx_staged_i16.cu", "gf3_rx_staged_f32.cu", "gf3_tx.cu", "gf3_sync.cu",
           "gf3_chan.cu", "gf3_stage.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
FLAGS += os.environ.get("GF3_EXTRA_FLAGS", "").split()
if "GF3_LIB_NAME" in os.environ:
    OBJDIR = os.path.join(HERE, "build_" + os.environ["GF3_LIB_NAME"].replace(".so", ""))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(OBJDIR, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):     # GPU box without a toolkit change: use the shipped build
            return LIB
        raise RuntimeError("nvcc not found at %s and no prebuilt %s" % (NVCC, LIB))

    def cc(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
