"""Builds csrc/*.cu into lrp_imagecaptioning_b200/liblrpcap.so for sm_100a (nvcc cross-compiles without a GPU).

The .so is kept in-tree (git-ignored) so it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "liblrpcap.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", CSRC,
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stamp(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False, debug=False):
    """debug=True: -DLRPCAP_DEBUG_BOUNDS (every epilogue address checked against its tensor's logical size, violation =
    printf + trap) into liblrpcap_dbg.so; select it at run time with LRPCAP_LIB=<path> (tools/run_bounds_check.py)."""
    global OBJ, LIB, NVCC_FLAGS
    if debug:
        OBJ = os.path.join(CSRC, "build_dbg")
        LIB = os.path.join(HERE, "liblrpcap_dbg.so")
        NVCC_FLAGS = NVCC_FLAGS + ["-DLRPCAP_DEBUG_BOUNDS"]
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "lrpcap.h")]
    stamp = _stamp(deps)
    stamp_file = os.path.join(OBJ, "stamp")
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            sys.stderr.write(r.stdout + r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True, debug="--debug" in sys.argv))
