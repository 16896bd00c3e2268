"""Generates tests/golden/*.npz — run in the build container, where /root/reference exists.

The reference keeps no golden vectors for this path (src/simd_test.cxx stores nothing and never
fails), so the fixtures are produced here from the reference's OWN kernels: the AVX2 / AVX-512
PhantomGrapeQuad code compiled verbatim from /root/reference/src/phantomquad_for_p3t_x86.hpp
(oracle/_ref, recipe in oracle/Makefile) on the input set of src/simd_test.cxx, next to the fp64
restatement's output on the same inputs; and, for the changeover correction, from the reference's
ChangeOver class and pair function compiled in oracle/_ref.  The GPU box has no /root/reference: its tests read
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
    # the reference's own fp64 NoSimd functors (src/soft_force.hpp:10-236, compiled from the reference source by
    # oracle/Makefile into oracle/_ref/libpetar_ref_nosimd.so): the oracle must reproduce these bit for bit
    out["nosimd_ep"] = ob.ref_nosimd("epep", epi, epj, P["eps"], P["r_out"], P["G"])
    out["nosimd_sp"] = ob.ref_nosimd("epsp_quad", epi, spj, P["eps"], G=P["G"])
    out["nosimd_sp_mono"] = ob.ref_nosimd("epsp_mono", epi, spj, P["eps"], G=P["G"])
    out["nosimd_nb"] = ob.ref_nosimd("search", epi, epj)
    out["nosimd_pp"] = ob.ref_nosimd("pp", epi, epj, G=P["G"])
    out["oracle_pp"] = ob.force_pp(epi, epj, P["G"])
    # a second parameter set with eps > 0 and G != 1 (the simd_test recipe has eps = 1e-4, G = 1)
    out["nosimd_ep_b"] = ob.ref_nosimd("epep", epi, epj, 3e-3, 2e-2, 0.37)
    out["nosimd_sp_b"] = ob.ref_nosimd("epsp_quad", epi, spj, 3e-3, G=0.37)
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

    # changeover correction: random pairs of every neighbour kind through the REFERENCE's own
    # calcAccPotShortWithLinearCutoff (src/hard.hpp:1408-1476, compiled in oracle/_ref from the reference
    # sources together with src/changeover.hpp), both the USE_GPU (float replay) and the all-double branch
    from petar_b200.types import PtclCorr
    rng = np.random.default_rng(2024)
    n, r_out_g, G = 1000, 2e-3, 0.7
    pi, pj = np.zeros(n, PtclCorr), np.zeros(n, PtclCorr)
    pi["pos"] = rng.uniform(-1, 1, (n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    pj["pos"] = pi["pos"] + (10 ** rng.uniform(-5, -2, n))[:, None] * d
    for q in (pi, pj):
        f = rng.uniform(1, 3, n)
        q["mass"], q["r_in"], q["r_out"], q["id"] = 10 ** rng.uniform(-7, -4, n), r_out_g / 10 * f, r_out_g * f, rng.integers(1, 1 << 40, n)
    kind = np.arange(n) % 3
    pj["status"][kind == 1] = -5.0; pj["mass_backup"][kind == 1] = pj["mass"][kind == 1]; pj["mass"][kind == 1] = 0.0
    pj["status"][kind == 2] = 3.0
    pi["acc"], pi["pot_tot"], pi["pot_soft"] = rng.normal(size=(n, 3)), rng.normal(size=n), rng.normal(size=n)
    pi["status"] = 1.0          # the pair function does not read it; keeps the one-particle loop's self-potential term out
    eps = np.where(np.arange(n) % 2 == 0, 0.0, 1e-4)
    out3 = dict(pi=pi, pj=pj, eps=eps, r_out=np.array(r_out_g), G=np.array(G))
    for replay in (0, 1):
        res = pi.copy()
        for k in range(n):
            ob.ref_changeover_pair(res[k:k + 1], pj[k:k + 1], float(eps[k]), r_out_g, G, replay)
        out3["ref_replay_fp32" if replay else "ref_fp64"] = res
    w = np.array([[ri, ro, dr, *ob.ref_changeover_w(ri, ro, dr)] for ri, ro, dr in
                  zip(10 ** rng.uniform(-5, -2, 500), 10 ** rng.uniform(-1.9, -1, 500), 10 ** rng.uniform(-6, -1, 500))])
    out3["w_table"] = w                                   # r_in, r_out, dr, calcAcc0W, calcPotW
    np.savez_compressed(os.path.join(OUT, "changeover_pairs.npz"), **out3)
    for f in ("simdtest.npz", "plummer1k_walks.npz", "changeover_pairs.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
