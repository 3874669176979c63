"""Build libregnet_b200.so in-tree with nvcc for sm_100a (no torch involved: the library is a plain C ABI).

    python -m regnet_for_3d_grasping_b200.build [--force]

The .so lands next to this file so that it travels with the repository snapshot to the GPU box.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libregnet_b200.so")
SOURCES = ["capi.cu", "fps.cu", "neighbors.cu", "gather.cu", "gemm_simt.cu", "gemm_tc.cu", "scorenet.cu", "region.cu", "grid.cu", "sa0_front.cu", "sa0_chain.cu", "train_ops.cu", "conv_train.cu", "generic_ops.cu", "train_gather.cu", "gemm_fused_a.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v", "--expt-relaxed-constexpr", "-Xfatbin", "-compress-all"]


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), "include", "regnet_b200.h"))
    return files


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ_DIR, src + ".o")
    log = os.path.join(OBJ_DIR, src + ".ptxas.txt")
    cmd = ["nvcc"] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout[-6000:]}")
    return obj


def build(force=False, verbose=False):
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _stale(LIB, _deps()):
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(_compile, sources))
    cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout[-4000:])
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
