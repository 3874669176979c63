"""Build recipe for oracle/_ref: the reference's own pn2_ext CUDA extension, kernel bodies untouched.

TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is imported by the product package.

The reference extension (multi_model/utils/pn2_utils/csrc/*.cu, main.cpp; built by
multi_model/utils/pn2_utils/setup.py:4-24 with `nvcc -O2`) targets torch 1.8 and no longer compiles
against torch 2.11 (THC removed, Tensor::type() dispatch removed).  This script applies a purely
mechanical API patch to a scratch copy under /tmp -- reference SOURCES are never
copied into this repository -- and compiles it for sm_100a into oracle/_ref/pn2_ext_ref.so.

The result is a torch extension module named `pn2_ext_ref` exporting the reference's 7 functions.  It needs
a GPU to run, so it is used (a) on the GPU box to generate golden vectors from the real reference kernels
(oracle/gen_golden_gpu.py) and (b) as the "reference CUDA kernels on B200" side baseline in bench.py.

Only runs where /root/reference exists (the build container).  On the GPU box the prebuilt .so travels.
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/multi_model/utils/pn2_utils/csrc"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "pn2_ext_ref.so")

THC_SHIM = r"""
#pragma once
// shim for the removed THC/THC.h: only the 5 macros the reference uses
#include <c10/cuda/CUDAException.h>
#include <c10/util/Exception.h>
#define THCudaCheck(x) C10_CUDA_CHECK(x)
#define THArgCheck(c, n, m) TORCH_CHECK(c, m)
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) TORCH_CHECK((a) == (b), #a " does not equal to " #b)
#endif
#ifndef CHECK_GT
#define CHECK_GT(a, b) TORCH_CHECK((a) > (b), #a " is not greater than " #b)
#endif
#ifndef CHECK_GE
#define CHECK_GE(a, b) TORCH_CHECK((a) >= (b), #a " is not greater or equal than " #b)
#endif
"""

SED_RULES = [
    (r"(\w+)\.type\(\)\.is_cuda\(\)", r"\1.is_cuda()"),
    (r"AT_DISPATCH_FLOATING_TYPES\((\w+)\.type\(\)", r"AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()"),
    (r"\.data<([^>]+)>\(\)", r".data_ptr<\1>()"),
    (r"(\w+)\.type\(\)\.toScalarType\(at::kLong\)", r"\1.options().dtype(at::kLong)"),
    (r", (\w+)\.type\(\)\)", r", \1.options())"),
]


def ref_available():
    return os.path.isdir(REF_CSRC)


def build(force=False, verbose=True):
    if os.path.exists(OUT_SO) and not force:
        return OUT_SO
    if not ref_available():
        raise RuntimeError("reference sources not present; oracle/_ref can only be built in the build container")
    import torch  # noqa: F401
    from torch.utils.cpp_extension import include_paths

    os.makedirs(OUT_DIR, exist_ok=True)
    work = tempfile.mkdtemp(prefix="pn2_ref_build_")
    try:
        os.makedirs(os.path.join(work, "THC"))
        with open(os.path.join(work, "THC", "THC.h"), "w") as f:
            f.write(THC_SHIM)
        srcs = []
        for name in sorted(os.listdir(REF_CSRC)):
            text = open(os.path.join(REF_CSRC, name)).read()
            if name.endswith(".cu"):
                for pat, rep in SED_RULES:
                    text = re.sub(pat, rep, text)
                if "THC/THC.h" not in text:
                    text = "#include <THC/THC.h>\n" + text
                srcs.append(name)
            elif name.endswith(".cpp"):
                srcs.append(name)
            with open(os.path.join(work, name), "w") as f:
                f.write(text)
        incs = include_paths("cuda") + [sysconfig.get_paths()["include"], work]
        inc_flags = [f"-I{p}" for p in incs]
        torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
        common = ["-DTORCH_EXTENSION_NAME=pn2_ext_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
                  "-D_GLIBCXX_USE_CXX11_ABI=1", "-std=c++17"]
        objs = []
        for s in srcs:
            obj = os.path.join(work, s + ".o")
            if s.endswith(".cu"):
                # reference flags: setup.py:4-5 -> nvcc -O2 (fmad left at its default = true)
                cmd = ["nvcc", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                       "-Xcompiler", "-fPIC", "-c", os.path.join(work, s), "-o", obj] + common + inc_flags
            else:
                cmd = ["g++", "-O2", "-fPIC", "-c", os.path.join(work, s), "-o", obj] + common + inc_flags
            if verbose:
                print("[build_ref]", " ".join(cmd[:6]), "...", s, flush=True)
            subprocess.check_call(cmd)
            objs.append(obj)
        cmd = ["g++", "-shared", "-o", OUT_SO] + objs + [
            f"-L{torch_lib}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
            "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{torch_lib}"]
        subprocess.check_call(cmd)
        # keep the SASS of the three search kernels' distance formula next to the binary (evidence for the
        # fmaf ordering the oracle pins; see oracle/pn2_oracle.c header)
        try:
            sass = subprocess.check_output(["cuobjdump", "-sass", OUT_SO], text=True)
            with open(os.path.join(OUT_DIR, "pn2_ext_ref.sass.txt"), "w") as f:
                f.write(sass)
        except Exception as e:  # pragma: no cover
            print("[build_ref] cuobjdump failed:", e)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return OUT_SO


def load():
    """Import the built extension as module `pn2_ext_ref` (needs torch imported first)."""
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(OUT_SO):
        raise ImportError("oracle/_ref/pn2_ext_ref.so not built")
    spec = importlib.util.spec_from_file_location("pn2_ext_ref", OUT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
