"""Builds ddnerf_b200/libddnerf_b200.so (the C-ABI library of include/ddnerf_b200.h) with nvcc
for sm_100a.  In-tree, so the .so travels to the GPU box with the repo snapshot.

    python -m ddnerf_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libddnerf_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
BASE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# Per-file extra flags.  The per-ray elementwise/scan kernels are compiled without FMA contraction
# so their rounding follows the reference's op-by-op fp32 arithmetic; the GEMM files keep it.
SOURCES = {
    "api.cu": [],
    "composite.cu": [],
    "sampler.cu": ["-fmad=false"],
    "encode.cu": ["-fmad=false"],
    "raygen.cu": ["-fmad=false"],
    "frame.cu": ["-fmad=false"],
    "raystore.cu": [],
    "dploss.cu": ["-fmad=false"],
    "mlp_f32.cu": [],
    "mlp_tc.cu": [],
    "mlp_tc_dw.cu": [],
    "tc_selftest.cu": [],
    "train_tail.cu": [],
}


def _stamp():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if not os.path.isfile(os.path.join(root, name)):
                continue
            with open(os.path.join(root, name), "rb") as f:
                h.update(name.encode() + f.read())
    h.update(repr((BASE_FLAGS, SOURCES)).encode())
    return h.hexdigest()


def _compile(src, extra):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [NVCC] + BASE_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    srcs = {s: f for s, f in SOURCES.items() if os.path.exists(os.path.join(CSRC, s))}
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda kv: _compile(*kv), srcs.items()))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if verbose:
        print(f"built {LIB} from {len(objs)} objects")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
