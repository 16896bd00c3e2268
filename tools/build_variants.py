"""Kernel tuning experiments: builds libpetar_b200.so variants with other compile-time switches of
pb_kernels.cu into petar_b200/lib/variants/<name>/ (load one with PETAR_B200_LIB=<path>).

    python tools/build_variants.py                      # the default A/B set below
    python tools/build_variants.py name:-DX=1,-DY=2 ... # explicit variants
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from petar_b200 import build as B

DEFAULT = {
    "notma": ["-DPB_TMA_IDS=0"],
}


def build_variant(name, defs):
    objdir = os.path.join(B.HERE, "_build")
    d = os.path.join(B.LIB, "variants", name)
    os.makedirs(d, exist_ok=True)
    obj = os.path.join(objdir, f"pb_kernels_{name}.o")
    subprocess.run([B._nvcc(), *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *defs, "-I", B.INC, "-I", B.CSRC,
                    "-c", "-o", obj, os.path.join(B.CSRC, "pb_kernels.cu")], check=True)
    subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xlinker", "-soname=libpetar_b200.so", "-o",
                    os.path.join(d, "libpetar_b200.so"), obj, os.path.join(objdir, "pb_engine.o"), os.path.join(objdir, "pb_walk.o"), os.path.join(objdir, "pb_corr.o"), "-lgomp"], check=True)
    print("built", d)


if __name__ == "__main__":
    B.build_engine()
    variants = dict(DEFAULT)
    if len(sys.argv) > 1:
        variants = {a.split(":", 1)[0]: a.split(":", 1)[1].split(",") for a in sys.argv[1:]}
    for name, defs in variants.items():
        build_variant(name, defs)
