"""CPU tests of the checker itself: the fp64 oracle against the reference's own kernels (golden
fixtures generated from oracle/_ref, and oracle/_ref live where it exists) and against analytic
known answers."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from petar_b200.types import EPISoft, EPJSoft, SPJQuad, ForceSoft
from conftest import GOLDEN

P = ob.SIMDTEST_PARAMS


def _rel_acc(a, b):
    return np.abs(a["acc"] - b["acc"]).max(axis=1) / np.linalg.norm(b["acc"], axis=1)


def _mk_epi(pos, rs):
    e = np.zeros(len(pos), dtype=EPISoft)
    e["pos"] = pos
    e["r_search"] = rs
    e["type"] = 1
    e["id"] = np.arange(len(pos)) + 1
    return e


def _mk_epj(pos, mass, rs):
    e = np.zeros(len(pos), dtype=EPJSoft)
    e["pos"] = pos
    e["mass"] = mass
    e["r_search"] = rs
    e["id"] = np.arange(len(pos)) + 1
    return e


# ---- golden vectors -------------------------------------------------------------------------
def test_mt19937_known_answer():
    """10000th output of mt19937 seeded with 5489 is 4123659995 (C++11 [rand.predef])."""
    import ctypes as C
    L = ob.oracle_lib()
    st = (C.c_uint32 * 625)()
    L.orc_mt_init(st, 5489)
    v = 0
    for _ in range(10000):
        v = L.orc_mt_int32(st)
    assert v == 4123659995


def test_simdtest_inputs_reproduce_golden():
    g = np.load(os.path.join(GOLDEN, "simdtest.npz"))
    epi, epj, spj = ob.simdtest_inputs()
    assert np.array_equal(epi["pos"][:4], g["epi_pos_head"])
    assert np.array_equal(spj[:2], g["spj_head"])
    # recipe facts of src/simd_test.cxx: N=2000 equal masses, r_search = r_out = 0.01 for everyone
    assert np.all(epj["mass"] == 1.0 / 2000) and np.all(epi["r_search"] == 0.01)
    assert 1.0 <= spj["pos"].min() and spj["pos"].max() <= 11.0


def test_oracle_matches_golden_bitwise():
    g = np.load(os.path.join(GOLDEN, "simdtest.npz"))
    epi, epj, spj = ob.simdtest_inputs()
    ep = ob.force_epep(epi, epj, P["eps"], P["r_out"], P["G"])
    sp = ob.force_epsp_quad(epi, spj, P["eps"], P["G"])
    for k in ("acc", "pot", "n_ngb"):
        assert np.array_equal(ep[k], g["oracle_ep"][k])
        assert np.array_equal(sp[k], g["oracle_sp"][k])


def test_oracle_equals_reference_nosimd_functors_bitwise():
    """THE PIN: the fp64 restatement (oracle_soft_force.c) reproduces, bit for bit, the outputs of the reference's own
    NoSimd functors — SearchNeighborEpEpNoSimd, CalcForceEpEpWithLinearCutoffNoSimd, CalcForceEpSpMonoNoSimd,
    CalcForceEpSpQuadNoSimd, CalcForcePPNoSimd (src/soft_force.hpp:10-236) — compiled from the reference source
    (oracle/ref_nosimd.cpp, -O2 -ffp-contract=off) and stored in tests/golden/simdtest.npz on the simd_test.cxx
    inputs, for the test's own parameters and for a second set with a large eps and G != 1."""
    g = np.load(os.path.join(GOLDEN, "simdtest.npz"))
    epi, epj, spj = ob.simdtest_inputs()
    assert ob.force_epep(epi, epj, P["eps"], P["r_out"], P["G"]).tobytes() == g["nosimd_ep"].tobytes()
    assert ob.force_epsp_quad(epi, spj, P["eps"], P["G"]).tobytes() == g["nosimd_sp"].tobytes()
    assert ob.force_epsp_mono(epi, spj, P["eps"], P["G"]).tobytes() == g["nosimd_sp_mono"].tobytes()
    assert ob.search_neighbor(epi, epj).tobytes() == g["nosimd_nb"].tobytes()
    assert ob.force_pp(epi, epj, P["G"]).tobytes() == g["nosimd_pp"].tobytes()
    assert ob.force_epep(epi, epj, 3e-3, 2e-2, 0.37).tobytes() == g["nosimd_ep_b"].tobytes()
    assert ob.force_epsp_quad(epi, spj, 3e-3, 0.37).tobytes() == g["nosimd_sp_b"].tobytes()


@pytest.mark.skipif(not ob.ref_nosimd_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_nosimd_live_equals_golden_and_walk_driver():
    """Where the reference is compiled (this container): its live outputs equal the committed vectors, the functors
    ACCUMULATE into acc / pot as the oracle does, and the oracle's multiwalk driver (index gather + EP functor + SP
    functor per walk) equals the reference functors applied walk by walk to the gathered arrays, bit for bit."""
    from petar_b200 import harness as hz
    g = np.load(os.path.join(GOLDEN, "simdtest.npz"))
    epi, epj, spj = ob.simdtest_inputs()
    f = ob.ref_nosimd("epep", epi, epj, P["eps"], P["r_out"], P["G"])
    assert f.tobytes() == g["nosimd_ep"].tobytes()
    f2 = ob.ref_nosimd("epsp_quad", epi, spj, P["eps"], G=P["G"], force=f.copy())        # accumulates onto the EP result
    o2 = ob.force_epsp_quad(epi, spj, P["eps"], P["G"], force=ob.force_epep(epi, epj, P["eps"], P["r_out"], P["G"]))
    assert f2.tobytes() == o2.tobytes()
    batch, _, prm, _ = hz.plummer_case(1000)
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    for w in range(batch.n_walk):
        i0, i1 = batch.i_off[w], batch.i_off[w + 1]
        je = np.ascontiguousarray(batch.epj[batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]]])
        js = np.ascontiguousarray(batch.spj[batch.id_spj[batch.sj_off[w]:batch.sj_off[w + 1]]])
        ei = np.ascontiguousarray(batch.epi[i0:i1])
        fw = ob.ref_nosimd("epep", ei, je, prm["eps"], prm["r_out"], prm["G"])
        fw = ob.ref_nosimd("epsp_quad", ei, js, prm["eps"], G=prm["G"], force=fw)
        assert fw.tobytes() == ref[i0:i1].tobytes(), w


@pytest.mark.parametrize("isa", ["avx2", "avx512"])
def test_oracle_vs_reference_simd_golden(isa):
    """CONTEXT ONLY (the pin is the bitwise NoSimd test above): the fp64 restatement agrees with the reference's own
    fp32 SIMD kernels to the SIMD kernels' precision: EP-EP (rsqrt + one Newton step) ~1e-6, EP-SP (raw rsqrt, no Newton step) within the
    reference test's own 7e-3 print threshold (src/simd_test.cxx:68); neighbour counts identical."""
    g = np.load(os.path.join(GOLDEN, "simdtest.npz"))
    ep, sp = g["oracle_ep"], g["oracle_sp"]
    rep, rsp = g[f"ref_{isa}_ep"], g[f"ref_{isa}_sp"]
    assert _rel_acc(rep, ep).max() < 2e-5 and np.median(_rel_acc(rep, ep)) < 2e-6
    assert np.abs((rep["pot"] - ep["pot"]) / ep["pot"]).max() < 1e-5
    assert _rel_acc(rsp, sp).max() < 7e-3
    assert np.abs((rsp["pot"] - sp["pot"]) / sp["pot"]).max() < 1e-2
    assert np.array_equal(rep["n_ngb"], ep["n_ngb"])
    assert np.array_equal(g[f"ref_{isa}_nb"]["n_ngb"], g["oracle_nb"]["n_ngb"])


@pytest.mark.skipif(not ob.ref_available("avx2"), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_simd_live_equals_golden():
    g = np.load(os.path.join(GOLDEN, "simdtest.npz"))
    epi, epj, spj = ob.simdtest_inputs()
    rep = ob.ref_force_epep(epi, epj, P["eps"], P["r_out"], P["G"], isa="avx2")
    assert np.allclose(rep["acc"], g["ref_avx2_ep"]["acc"], rtol=0, atol=0)
    assert np.array_equal(rep["n_ngb"], g["ref_avx2_ep"]["n_ngb"])


def test_plummer1k_walks_golden():
    from petar_b200 import harness as hz
    g = np.load(os.path.join(GOLDEN, "plummer1k_walks.npz"))
    batch, _, prm, _ = hz.plummer_case(1000)
    assert batch.n_walk == int(g["n_walk"]) and np.array_equal(batch.ej_off, g["ej_off"])
    f = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    assert np.array_equal(f["n_ngb"], g["oracle"]["n_ngb"])
    assert np.allclose(f["acc"], g["oracle"]["acc"], rtol=1e-13, atol=0)
    # reference SIMD path on the same lists (all masses > 0, all types 1: no filtering differences)
    for isa in ("avx2", "avx512"):
        r = g[f"ref_{isa}"]
        assert np.array_equal(r["n_ngb"], f["n_ngb"])
        assert np.median(_rel_acc(r, f)) < 1e-4      # dominated by the no-Newton-step SP kernel


# ---- analytic known answers -----------------------------------------------------------------
def test_two_body_outside_cutoff():
    epi = _mk_epi([[0.0, 0, 0]], 0.1)
    epj = _mk_epj([[3.0, 4.0, 0.0]], [2.0], 0.1)
    f = ob.force_epep(epi, epj, 0.0, 0.5, 1.5)
    assert np.allclose(f["acc"][0], 1.5 * 2.0 / 125.0 * np.array([3.0, 4.0, 0.0]), rtol=1e-15)
    assert np.isclose(f["pot"][0], -1.5 * 2.0 / 5.0, rtol=1e-15)
    assert f["n_ngb"][0] == 0


def test_self_only_list():
    """i in its own list: zero force, pot = -G m / r_out (clamp), counted as neighbour (SURVEY §7 hard part 2)."""
    epi = _mk_epi([[0.3, -0.2, 0.9]], 0.02)
    epj = _mk_epj([[0.3, -0.2, 0.9]], [1e-3], 0.02)
    f = ob.force_epep(epi, epj, 0.0, 0.01, 1.0)
    assert np.all(f["acc"][0] == 0.0)
    assert np.isclose(f["pot"][0], -1e-3 / 0.01, rtol=1e-15)
    assert f["n_ngb"][0] == 1


def test_inside_cutoff_is_linear():
    """inside r_out the force is G m dx / r_out^3 (hence 'linear cutoff')."""
    epi = _mk_epi([[0.0, 0, 0]], 0.0)
    for d in (1e-4, 3e-3, 9.9e-3):
        epj = _mk_epj([[d, 0.0, 0.0]], [1.0], 0.0)
        f = ob.force_epep(epi, epj, 0.0, 0.01, 1.0)
        assert np.isclose(f["acc"][0, 0], d / 0.01 ** 3, rtol=1e-13)
        assert np.isclose(f["pot"][0], -1.0 / 0.01, rtol=1e-13)


def test_neighbour_tie_is_strict():
    """r2 < rs^2 is strict: a j exactly at max(rs_i, rs_j) is not a neighbour; eps is not in the test."""
    epi = _mk_epi([[0.0, 0, 0]], 0.25)
    epj = _mk_epj([[0.5, 0.0, 0.0], [0.4999999, 0, 0]], [1.0, 1.0], [0.5, 0.5])
    f = ob.force_epep(epi, epj, 10.0, 0.01, 1.0)
    assert f["n_ngb"][0] == 1


def test_zero_mass_j_counts_but_no_force():
    epi = _mk_epi([[0.0, 0, 0]], 0.1)
    epj = _mk_epj([[0.05, 0.0, 0.0]], [0.0], 0.1)
    f = ob.force_epep(epi, epj, 0.0, 0.01, 1.0)
    assert np.all(f["acc"] == 0) and f["pot"][0] == 0 and f["n_ngb"][0] == 1


def test_eps_larger_than_rout_makes_clamp_inert():
    rng = np.random.default_rng(1)
    epi = _mk_epi(rng.normal(size=(8, 3)), 0.0)
    epj = _mk_epj(rng.normal(size=(50, 3)), rng.random(50), 0.0)
    a = ob.force_epep(epi, epj, 0.3, 0.2, 1.0)
    b = ob.force_epep(epi, epj, 0.3, 0.0, 1.0)
    assert np.array_equal(a["acc"], b["acc"]) and np.array_equal(a["pot"], b["pot"])


def test_quad_with_zero_quadrupole_equals_monopole():
    rng = np.random.default_rng(2)
    epi = _mk_epi(rng.normal(size=(16, 3)), 0.0)
    spj = np.zeros(40, dtype=SPJQuad)
    spj["mass"] = rng.random(40)
    spj["pos"] = rng.normal(size=(40, 3)) + 5.0
    q = ob.force_epsp_quad(epi, spj, 0.01, 1.0)
    m = ob.force_epsp_mono(epi, spj, 0.01, 1.0)
    assert np.allclose(q["acc"], m["acc"], rtol=1e-14) and np.allclose(q["pot"], m["pot"], rtol=1e-14)


def test_pure_trace_quadrupole_is_radial_only():
    """Q = s*I: traceless part vanishes, so (with eps = 0) the quadrupole terms cancel exactly:
    A = m r^-3 - 1.5*3s r^-5 + 7.5 s r^2 r^-7 and B qr = -3 s r^-5 dx  =>  force = monopole."""
    epi = _mk_epi([[0.0, 0, 0]], 0.0)
    spj = np.zeros(1, dtype=SPJQuad)
    spj["mass"] = 2.0
    spj["pos"] = [[1.0, 2.0, 2.0]]
    spj["quad"] = [[0.7, 0.7, 0.7, 0, 0, 0]]
    q = ob.force_epsp_quad(epi, spj, 0.0, 1.0)
    m = ob.force_epsp_mono(epi, spj, 0.0, 1.0)
    assert np.allclose(q["acc"], m["acc"], rtol=1e-13) and np.allclose(q["pot"], m["pot"], rtol=1e-13)


def test_quadrupole_of_two_points_converges_to_direct_sum():
    """A superparticle built from two point masses reproduces their direct sum to O((d/r)^3)."""
    x1, x2, m1, m2 = np.array([0.01, 0.0, 0.0]), np.array([-0.01, 0.005, 0.0]), 1.0, 1.0
    cm = (m1 * x1 + m2 * x2) / (m1 + m2)
    qq = np.zeros((3, 3))
    for x, m in ((x1, m1), (x2, m2)):
        d = x - cm
        qq += m * np.outer(d, d)
    spj = np.zeros(1, dtype=SPJQuad)
    spj["mass"] = m1 + m2
    spj["pos"] = [cm]
    spj["quad"] = [[qq[0, 0], qq[1, 1], qq[2, 2], qq[0, 1], qq[0, 2], qq[1, 2]]]
    epi = _mk_epi([[1.0, 0.5, -0.3]], 0.0)
    epj = _mk_epj([x1, x2], [m1, m2], 0.0)
    direct = ob.force_pp(epi, epj, 1.0)
    quad = ob.force_epsp_quad(epi, spj, 0.0, 1.0)
    mono = ob.force_epsp_mono(epi, spj, 0.0, 1.0)
    eq = np.abs(quad["acc"] - direct["acc"]).max()
    em = np.abs(mono["acc"] - direct["acc"]).max()
    assert eq < 1e-5 and eq < 0.05 * em


def test_empty_lists():
    epi = _mk_epi([[0.0, 0, 0]], 0.1)
    f = ob.force_epep(epi, _mk_epj(np.zeros((0, 3)), [], []), 0.0, 0.01, 1.0)
    assert np.all(f["acc"] == 0) and f["pot"][0] == 0 and f["n_ngb"][0] == 0
    f = ob.force_epsp_quad(epi, np.zeros(0, dtype=SPJQuad), 0.0, 1.0)
    assert np.all(f["acc"] == 0)


# ---- changeover correction: the restatement against the reference's own function ----------------
needs_ref_co = pytest.mark.skipif(not ob.ref_changeover_available(), reason="oracle/_ref changeover libraries not built")


@needs_ref_co
def test_changeover_functions_bit_exact_vs_reference():
    rng = np.random.default_rng(1)
    for _ in range(5000):
        r_in = 10 ** rng.uniform(-5, -2); r_out = r_in * rng.uniform(2, 20); dr = 10 ** rng.uniform(-6, -1)
        assert ob.changeover_w(r_in, r_out, dr) == ob.ref_changeover_w(r_in, r_out, dr)
    # the three regimes: inside r_in, the polynomial, beyond r_out
    a0, p0 = ob.changeover_w(1e-3, 1e-2, 5e-4); a1, p1 = ob.changeover_w(1e-3, 1e-2, 2e-2)
    assert a0 == 1.0 and p0 == 1.0 and a1 == 0.0 and p1 == (1.0 + 9.0 / 11.0) / 1e-2 * 2e-2


@needs_ref_co
@pytest.mark.parametrize("replay_fp32", [False, True])
def test_changeover_pair_bit_exact_vs_reference(replay_fp32):
    """oracle_changeover.c against calcAccPotShortWithLinearCutoff (reference src/hard.hpp:1408-1476) as
    compiled in oracle/_ref from the reference sources, USE_GPU and P3T_64BIT branches."""
    from petar_b200.types import PtclCorr
    rng = np.random.default_rng(2)
    r_out_g = 2e-3
    for t in range(6000):
        pi, pj = np.zeros(1, PtclCorr), np.zeros(1, PtclCorr)
        pi["pos"] = rng.uniform(-1, 1, 3)
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        pj["pos"] = pi["pos"] + 10 ** rng.uniform(-5, -2) * d
        for q in (pi, pj):
            f = rng.uniform(1, 3)
            q["mass"], q["r_in"], q["r_out"] = 10 ** rng.uniform(-7, -4), r_out_g / 10 * f, r_out_g * f
        if t % 3 == 1:
            pj["status"], pj["mass_backup"], pj["mass"] = -5.0, pj["mass"], 0.0
        if t % 3 == 2:
            pj["status"] = 3.0
        pi["acc"], pi["pot_tot"], pi["pot_soft"] = rng.normal(size=3), rng.normal(), rng.normal()
        a, b = pi.copy(), pi.copy()
        eps = (0.0, 1e-4)[t % 2]
        ob.changeover_pair(a, pj, eps, r_out_g, 0.7, replay_fp32)
        ob.ref_changeover_pair(b, pj, eps, r_out_g, 0.7, replay_fp32)
        assert a.tobytes() == b.tobytes()


def test_changeover_neighbor_loop_matches_pairwise_application():
    from petar_b200.types import PtclCorr, LARGE_FLOAT
    from petar_b200 import harness
    P = harness.kroupa_binary_particles(3000, f_bin=0.2)
    p = harness.corr_particles(P)
    off, idx = harness.neighbor_lists(P["pos"], P["rs"])
    assert (p["status"] < 0).any() and (p["status"] > 0).any() and np.diff(off).max() > 3
    out = ob.correct_force_tree_neighbor(p.copy(), off, idx, p, 0.0, P["prm"]["r_out"], 1.0, False)
    for i in np.random.default_rng(0).choice(len(p), 200, replace=False):
        q = p[i:i + 1].copy()
        if q["status"][0] == 0.0 and q["mass_backup"][0] == 0.0:
            q["pot_tot"] += q["mass"] / P["prm"]["r_out"]; q["pot_soft"] += q["mass"] / P["prm"]["r_out"]
        for j in idx[off[i]:off[i + 1]]:
            if p["id"][j] != q["id"][0]:
                ob.changeover_pair(q, p[j:j + 1].copy(), 0.0, P["prm"]["r_out"], 1.0, False)
        assert q.tobytes() == out[i:i + 1].tobytes()


def test_changeover_oracle_vs_committed_reference_vectors():
    """tests/golden/changeover_pairs.npz: outputs of the reference's own pair function and ChangeOver class
    (generated by tests/golden/make_golden.py where /root/reference exists) — bit for bit, no oracle/_ref needed."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "changeover_pairs.npz"))
    pi, pj, eps = g["pi"], g["pj"], g["eps"]
    for replay, key in ((0, "ref_fp64"), (1, "ref_replay_fp32")):
        out = pi.copy()
        for k in range(len(pi)):
            ob.changeover_pair(out[k:k + 1], pj[k:k + 1], float(eps[k]), float(g["r_out"]), float(g["G"]), replay)
        assert out.tobytes() == g[key].tobytes()
    for r_in, r_out, dr, a, p in g["w_table"]:
        assert ob.changeover_w(r_in, r_out, dr) == (a, p)


def test_changeover_correction_properties():
    """Physics of the pair correction (fp64 branch): nothing changes beyond both the changeover radius and the
    cutoff radius; inside r_in the pair's soft force is removed entirely, i.e. the correction only takes the
    kernel's clamped term back out (acc_after = acc_before + G m dr / r_out^3); the corrected pair force obeys
    Newton's third law; the part of the pair potential left to the hard integrator shrinks with separation."""
    from petar_b200.types import PtclCorr
    G, r_out_g = 1.0, 2e-3

    def pair(sep, mi=3e-6, mj=5e-6, fi=1.0, fj=1.5):
        pi, pj = np.zeros(1, PtclCorr), np.zeros(1, PtclCorr)
        pi["pos"], pj["pos"] = [0.1, -0.2, 0.3], [0.1 + sep, -0.2, 0.3]
        pi["mass"], pj["mass"] = mi, mj
        pi["r_in"], pi["r_out"], pj["r_in"], pj["r_out"] = 0.1 * r_out_g * fi, r_out_g * fi, 0.1 * r_out_g * fj, r_out_g * fj
        pi["id"], pj["id"] = 1, 2
        return pi, pj

    # beyond max(r_out_i, r_out_j) = 1.5 r_out and beyond the global cutoff: exactly nothing to correct in acc
    pi, pj = pair(2.0 * r_out_g)
    a = pi.copy(); ob.changeover_pair(a, pj, 0.0, r_out_g, G, 0)
    assert np.array_equal(a["acc"], pi["acc"]) and a["pot_tot"][0] == 0.0
    # inside r_in of the pair: k = 0, so the correction only takes the clamped (linear-cutoff) term back out
    sep = 0.05 * r_out_g
    pi, pj = pair(sep)
    a = pi.copy(); ob.changeover_pair(a, pj, 0.0, r_out_g, G, 0)
    assert np.allclose(a["acc"][0], [G * pj["mass"][0] * (-sep) / r_out_g ** 3, 0.0, 0.0], rtol=1e-11, atol=1e-300)
    # Newton's third law inside the changeover region
    pi, pj = pair(0.7 * r_out_g)
    a, b = pi.copy(), pj.copy()
    ob.changeover_pair(a, pj, 0.0, r_out_g, G, 0)
    ob.changeover_pair(b, pi, 0.0, r_out_g, G, 0)
    assert np.allclose(pi["mass"] * a["acc"], -pj["mass"] * b["acc"], rtol=1e-13, atol=0)
    # potential: tot - soft of the pair is the part the hard integrator owns, positive and shrinking with separation
    own = [(lambda a: (a["pot_soft"] - a["pot_tot"])[0])((lambda p: (ob.changeover_pair(p[0], p[1], 0.0, r_out_g, G, 0), p[0])[1])((pair(s)[0].copy(), pair(s)[1])))
           for s in (0.2 * r_out_g, 0.6 * r_out_g, 1.2 * r_out_g)]
    assert own[0] > own[1] > own[2] >= 0.0
