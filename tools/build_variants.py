"""Kernel tuning experiments: builds libpetar_b200.so variants with other inner-loop unroll factors
into petar_b200/lib/variants/<name>/ (load one with PETAR_B200_LIB=<path>)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from petar_b200 import build as B

B.build_engine()
objdir = os.path.join(B.HERE, "_build")
for eu, su in [(2, 2), (8, 2), (4, 1), (4, 4), (8, 4), (2, 4)]:
    d = os.path.join(B.LIB, "variants", f"ep{eu}_sp{su}")
    os.makedirs(d, exist_ok=True)
    obj = os.path.join(objdir, f"pb_kernels_ep{eu}_sp{su}.o")
    subprocess.run([B._nvcc(), *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")], f"-DPB_EP_UNROLL={eu}", f"-DPB_SP_UNROLL={su}", "-I", B.INC, "-I", B.CSRC,
                    "-c", "-o", obj, os.path.join(B.CSRC, "pb_kernels.cu")], check=True)
    subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xlinker", "-soname=libpetar_b200.so", "-o",
                    os.path.join(d, "libpetar_b200.so"), obj, os.path.join(objdir, "pb_engine.o"), os.path.join(objdir, "pb_walk.o"), "-lgomp"], check=True)
    print("built", d)
