"""Generates tests/golden/*.npz — run in the build container, where /root/reference exists.

The reference keeps no golden vectors for this path (src/simd_test.cxx stores nothing and never
fails), so the fixtures are produced here from the reference's OWN kernels: the AVX2 / AVX-512
PhantomGrapeQuad code compiled verbatim from /root/reference/src/phantomquad_for_p3t_x86.hpp
(oracle/_ref, recipe in oracle/Makefile) on the input set of src/simd_test.cxx, next to the fp64
restatement's output on the same inputs.  The GPU box has no /root/reference: its tests read
only the committed .npz files.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402
from petar_b200 import harness as hz  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    ob.build()
    P = ob.SIMDTEST_PARAMS
    epi, epj, spj = ob.simdtest_inputs()
    out = dict(
        input_sha256=np.array(digest(epi, epj, spj)),
        epi_pos_head=epi["pos"][:4].copy(), spj_head=spj[:2].copy(),
        oracle_ep=ob.force_epep(epi, epj, P["eps"], P["r_out"], P["G"]),
        oracle_sp=ob.force_epsp_quad(epi, spj, P["eps"], P["G"]),
        oracle_sp_mono=ob.force_epsp_mono(epi, spj, P["eps"], P["G"]),
        oracle_nb=ob.search_neighbor(epi, epj),
    )
    for isa in ("avx2", "avx512"):
        out[f"ref_{isa}_ep"] = ob.ref_force_epep(epi, epj, P["eps"], P["r_out"], P["G"], isa=isa)
        out[f"ref_{isa}_sp"] = ob.ref_force_epsp_quad(epi, spj, P["eps"], P["G"], isa=isa)
        out[f"ref_{isa}_nb"] = ob.ref_search_neighbor(epi, epj, isa=isa)
    np.savez_compressed(os.path.join(OUT, "simdtest.npz"), **out)

    # N = 1000 Plummer through the walk-list harness (theta 0.3, leaf 20, group 512), PeTar's
    # automatic parameters; fp64 oracle and reference-SIMD results per i-particle (walk-major)
    batch, epi_src, prm, _ = hz.plummer_case(1000)
    f64 = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    out2 = dict(
        input_sha256=np.array(digest(batch.epi, batch.epj, batch.spj, batch.id_epj, batch.id_spj)),
        n_walk=np.array(batch.n_walk), i_off=batch.i_off, ej_off=batch.ej_off, sj_off=batch.sj_off,
        r_out=np.array(prm["r_out"]), eps=np.array(prm["eps"]), G=np.array(prm["G"]),
        oracle=f64,
    )
    for isa in ("avx2", "avx512"):
        out2[f"ref_{isa}"], _ = ob.ref_walks_index(batch, prm["eps"], prm["r_out"], prm["G"], isa=isa)
    np.savez_compressed(os.path.join(OUT, "plummer1k_walks.npz"), **out2)
    for f in ("simdtest.npz", "plummer1k_walks.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
