"""Builds the in-tree native libraries with nvcc / g++ (sm_100a only).

    python -m petar_b200.build            # libpetar_b200.so (+ shim, harness)

Outputs go to ``petar_b200/lib/`` (git-ignored, but they travel to the GPU box with gpurun).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
INC = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp,-O3",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        sys.stdout.write(" ".join(cmd) + "\n" + out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd))
    return out.stdout + out.stderr


def build_engine(force=False, verbose=False):
    """libpetar_b200.so: CUDA kernels + C ABI."""
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libpetar_b200.so")
    # pb_walk.cu (tree walk, fp64 geometry) and pb_corr.cu (changeover correction, fp64) are compiled
    # without FMA contraction so that their results are bit-identical to host code; the force kernels
    # keep the default
    units = (("pb_kernels.cu", []), ("pb_engine.cu", []), ("pb_walk.cu", ["-fmad=false"]), ("pb_corr.cu", ["-fmad=false"]), ("pb_plan.cu", []), ("pb_kernels_ws.cu", []), ("pb_pack.cu", ["-fmad=false"]))
    srcs = [os.path.join(CSRC, f) for f, _ in units]
    deps = srcs + [os.path.join(CSRC, "pb_device.h"), os.path.join(CSRC, "pb_pairs.cuh"), os.path.join(INC, "petar_b200.h")]
    if force or _newer(target, deps):
        objdir = os.path.join(HERE, "_build")
        os.makedirs(objdir, exist_ok=True)
        log, objs = "", []
        for f, extra in units:
            obj = os.path.join(objdir, f.replace(".cu", ".o"))
            log += _run([_nvcc(), *NVCC_FLAGS, *extra, "-I", INC, "-I", CSRC, "-c", "-o", obj, os.path.join(CSRC, f)], verbose)
            objs.append(obj)
        log += _run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xlinker", "-soname=libpetar_b200.so",
                     "-o", target, *objs, "-lgomp"], verbose)
        with open(os.path.join(LIB, "ptxas_engine.log"), "w") as f:
            f.write(log)
    return target


def build_shim(force=False, verbose=False):
    """libpetar_b200_shim.so / libpetar_b200_shim_direct.so: the C++ translation unit that defines
    PeTar's functor symbols, compiled over the POD mirrors (host C++ only, no CUDA) in the index
    mode (-DPARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX) and in the non-index mode."""
    os.makedirs(LIB, exist_ok=True)
    src = os.path.join(CSRC, "force_gpu_b200.cpp")
    deps = [src, os.path.join(CSRC, "force_gpu_b200.hpp"), os.path.join(INC, "petar_b200.h"), os.path.join(INC, "petar_b200_types.h")]
    out = []
    for name, extra in (("libpetar_b200_shim.so", ["-DPARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX"]), ("libpetar_b200_shim_direct.so", [])):
        target = os.path.join(LIB, name)
        if force or _newer(target, deps):
            _run(["g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-fopenmp", "-shared", "-DPB_STANDALONE_MIRRORS", "-DUSE_GPU", "-DGPU_PROFILE", "-DUSE_QUAD",
                  *extra, "-I", INC, "-I", CSRC, "-o", target, src, "-L", LIB, "-lpetar_b200", "-Wl,-rpath,$ORIGIN"], verbose)
        out.append(target)
    return out


def build_harness(force=False, verbose=False):
    """libpetar_b200_harness.so: FDPS-like tree / walk-list generator for synthetic inputs (host C++)."""
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libpetar_b200_harness.so")
    src = os.path.join(HERE, "harness", "tree_walk.cpp")
    if not os.path.exists(src):
        return None
    deps = [src, os.path.join(INC, "petar_b200_types.h")]
    if force or _newer(target, deps):
        _run(["g++", "-O3", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-I", INC, "-o", target, src], verbose)
    return target


def build_all(force=False, verbose=False):
    return [build_engine(force, verbose), *build_shim(force, verbose), build_harness(force, verbose)]


if __name__ == "__main__":
    for t in build_all(force="--force" in sys.argv, verbose=True):
        print("built", t)
