"""GPU parity tests: the CUDA path, called through the C++ shim's PeTar functors and the C ABI,
against the fp64 oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star): acceleration and potential relative error vs the fp64
reference path <= 1e-6 median and <= 1e-4 max; neighbour counts bit-exact except for pairs within
fp32 rounding of the search radius, which are counted and reported.  Relative error of an
acceleration is |a_gpu - a_ref| / |a_ref| (vector norm)."""
import ctypes as C
import os

import numpy as np
import pytest

from petar_b200 import engine, harness as hz
from petar_b200.types import EPISoft, EPJSoft, SPJQuad, ForceSoft
from petar_b200.walks import WalkBatch
from oracle import binding as ob
from conftest import GOLDEN

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("coords0")]   # kernel-level parity: see conftest.coords0

TOL_MED = 1e-6
TOL_MAX = 1e-4
P = ob.SIMDTEST_PARAMS


def rel_err(f, ref):
    ea = np.linalg.norm(f["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ep = np.abs((f["pot"] - ref["pot"]) / ref["pot"])
    return ea, ep


def check_tol(f, ref, what, tol_med=TOL_MED, tol_max=TOL_MAX):
    ea, ep = rel_err(f, ref)
    print(f"[{what}] acc rel err median {np.median(ea):.3e} p99 {np.percentile(ea, 99):.3e} max {ea.max():.3e} | "
          f"pot median {np.median(ep):.3e} max {ep.max():.3e}")
    assert np.median(ea) <= tol_med and ea.max() <= tol_max, what
    assert np.median(ep) <= tol_med and ep.max() <= tol_max, what


def count_mismatch_report(batch, f, ref, what):
    """Neighbour counts must be identical except for pairs whose r^2 lies within fp32 rounding of
    max(rs_i, rs_j)^2; every mismatching i is checked to have such a borderline pair."""
    bad = np.nonzero(f["n_ngb"] != ref["n_ngb"])[0]
    n_border = 0
    for k in bad:
        w = np.searchsorted(batch.i_off, k, side="right") - 1
        e = batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]]
        d = batch.epj["pos"][e] - batch.epi["pos"][k]
        r2 = (d * d).sum(1)
        rs2 = np.maximum(batch.epj["r_search"][e], batch.epi["r_search"][k]) ** 2
        border = np.abs(r2 - rs2) <= 4e-6 * rs2
        assert border.sum() >= abs(int(f["n_ngb"][k]) - int(ref["n_ngb"][k])), f"{what}: count differs with no borderline pair (i={k})"
        n_border += 1
    print(f"[{what}] neighbour-count mismatches: {len(bad)} of {len(f)} (all borderline: {n_border})")
    return len(bad)


@pytest.fixture(autouse=True)
def _defaults():
    for k, v in (("coords", 0), ("streams", 2), ("jchunk", 0), ("nr", 0), ("cull", 1), ("min_slot_work", 0)):
        engine.set_option(k, v)
    yield


# ---- the reference's own test case ------------------------------------------------------------
def test_simdtest_case_like_reference():
    """reference src/simd_test.cxx:139-155: one walk, identity index lists, tag = 1."""
    epi, epj, spj = ob.simdtest_inputs()
    g = np.load(os.path.join(GOLDEN, "simdtest.npz"))
    batch = WalkBatch.single(epi, epj, spj)
    force_gpu = np.zeros(len(epi), dtype=ForceSoft)
    t = batch.pointer_tables(force_gpu)
    f_ep_ep_gpu = engine.CalcForceWithLinearCutoffCUDAMultiWalk(0, P["eps"] ** 2, P["r_out"] ** 2, P["G"])
    assert f_ep_ep_gpu(1, 1, t.epi_ptrs, t.n_epi, t.id_epj_ptrs, t.n_epj, t.id_spj_ptrs, t.n_spj, epj, len(epj), spj, len(spj), True) == 0
    assert f_ep_ep_gpu(1, 1, t.epi_ptrs, t.n_epi, t.id_epj_ptrs, t.n_epj, t.id_spj_ptrs, t.n_spj, epj, len(epj), spj, len(spj), False) == 0
    assert engine.RetrieveForceCUDA(1, 1, t.n_epi, t.force_ptrs) == 0
    # oracle: NoSimd EP-EP + NoSimd EP-SP(quad), as simd_test compares (force + force_sp vs force_gpu)
    ref = g["oracle_ep"].copy()
    ref["acc"] += g["oracle_sp"]["acc"]
    ref["pot"] += g["oracle_sp"]["pot"]
    check_tol(force_gpu, ref, "simd_test EP+SP vs fp64 oracle (golden)")
    assert np.array_equal(force_gpu["n_ngb"], ref["n_ngb"])
    # context: the reference's own AVX-512 path against the same oracle
    r = g["ref_avx512_ep"].copy()
    r["acc"] += g["ref_avx512_sp"]["acc"]
    r["pot"] += g["ref_avx512_sp"]["pot"]
    ea, ep = rel_err(r, ref)
    print(f"[context] reference AVX-512 path: acc median {np.median(ea):.3e} max {ea.max():.3e} pot max {ep.max():.3e}")


def test_ep_only_and_sp_only_split():
    epi, epj, spj = ob.simdtest_inputs()
    g = np.load(os.path.join(GOLDEN, "simdtest.npz"))
    f = engine.calc_force_all_and_write_back(WalkBatch.single(epi, epj, spj[:0]), P["eps"], P["r_out"], P["G"])
    check_tol(f, g["oracle_ep"], "EP only")
    assert np.array_equal(f["n_ngb"], g["oracle_ep"]["n_ngb"])
    f = engine.calc_force_all_and_write_back(WalkBatch.single(epi, epj[:0], spj), P["eps"], P["r_out"], P["G"])
    check_tol(f, g["oracle_sp"], "SP only")
    assert np.all(f["n_ngb"] == 0)


# ---- analytic known answers -------------------------------------------------------
def _epi(pos, rs):
    e = np.zeros(len(pos), dtype=EPISoft)
    e["pos"], e["r_search"], e["type"] = pos, rs, 1
    return e


def _epj(pos, mass, rs):
    e = np.zeros(len(pos), dtype=EPJSoft)
    e["pos"], e["mass"], e["r_search"] = pos, mass, rs
    return e


def test_known_answers():
    none = np.zeros(0, dtype=SPJQuad)
    # self-only list: acc = 0, pot = -G m / r_out, n_ngb = 1
    f = engine.calc_force_all_and_write_back(WalkBatch.single(_epi([[0.3, -0.2, 0.9]], 0.02), _epj([[0.3, -0.2, 0.9]], [1e-3], 0.02), none), 0.0, 0.01, 1.0)
    assert np.all(f["acc"][0] == 0.0) and np.isclose(f["pot"][0], -0.1, rtol=1e-6) and f["n_ngb"][0] == 1
    # two-body outside the cutoff
    f = engine.calc_force_all_and_write_back(WalkBatch.single(_epi([[0.0, 0, 0]], 0.1), _epj([[3.0, 4.0, 0.0]], [2.0], 0.1), none), 0.0, 0.5, 1.5)
    assert np.allclose(f["acc"][0], 1.5 * 2.0 / 125.0 * np.array([3.0, 4.0, 0.0]), rtol=2e-6)
    assert np.isclose(f["pot"][0], -0.6, rtol=1e-6) and f["n_ngb"][0] == 0
    # inside the cutoff the force is linear in dx
    for d in (1e-4, 3e-3, 9.9e-3):
        f = engine.calc_force_all_and_write_back(WalkBatch.single(_epi([[0.0, 0, 0]], 0.0), _epj([[d, 0.0, 0.0]], [1.0], 0.0), none), 0.0, 0.01, 1.0)
        assert np.isclose(f["acc"][0, 0], d / 0.01 ** 3, rtol=2e-6) and np.isclose(f["pot"][0], -100.0, rtol=1e-6)
    # zero-mass j: counted, no force
    f = engine.calc_force_all_and_write_back(WalkBatch.single(_epi([[0.0, 0, 0]], 0.1), _epj([[0.05, 0.0, 0.0]], [0.0], 0.1), none), 0.0, 0.01, 1.0)
    assert np.all(f["acc"] == 0) and f["pot"][0] == 0 and f["n_ngb"][0] == 1
    # strict '<' at the search radius (exactly representable numbers)
    f = engine.calc_force_all_and_write_back(WalkBatch.single(_epi([[0.0, 0, 0]], 0.25), _epj([[0.5, 0, 0], [0.25, 0, 0]], [1.0, 1.0], [0.5, 0.125]), none), 0.0, 0.01, 1.0)
    assert f["n_ngb"][0] == 0 + 0      # r = rs exactly for both: not neighbours
    # empty walk lists
    f = engine.calc_force_all_and_write_back(WalkBatch.single(_epi([[0.0, 0, 0]], 0.1), _epj(np.zeros((0, 3)), [], []), none), 0.0, 0.01, 1.0)
    assert np.all(f["acc"] == 0) and f["pot"][0] == 0 and f["n_ngb"][0] == 0


def test_quadrupole_known_answers():
    rng = np.random.default_rng(2)
    epi = _epi(rng.normal(size=(70, 3)), 0.0)
    none_e = _epj(np.zeros((0, 3)), [], [])
    spj = np.zeros(40, dtype=SPJQuad)
    spj["mass"] = rng.random(40)
    spj["pos"] = rng.normal(size=(40, 3)) + 6.0
    # zero quadrupole == monopole
    f = engine.calc_force_all_and_write_back(WalkBatch.single(epi, none_e, spj), 0.01, 0.0, 1.0)
    check_tol(f, ob.force_epsp_mono(epi, spj, 0.01, 1.0), "zero quadrupole vs monopole oracle")
    # pure-trace quadrupole == monopole (eps = 0)
    spj["quad"][:, 0:3] = rng.random(40)[:, None] * 0.05
    f = engine.calc_force_all_and_write_back(WalkBatch.single(epi, none_e, spj), 0.0, 0.0, 1.0)
    check_tol(f, ob.force_epsp_mono(epi, spj, 0.0, 1.0), "pure-trace quadrupole vs monopole oracle", tol_med=3e-6)


# ---- walk lists from the harness --------------------------------------------------------------
def test_plummer1k_walks_vs_golden():
    g = np.load(os.path.join(GOLDEN, "plummer1k_walks.npz"))
    batch, _, prm, _ = hz.plummer_case(1000)
    assert np.array_equal(batch.ej_off, g["ej_off"]), "harness did not reproduce the golden walk lists"
    f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    check_tol(f, g["oracle"], "Plummer N=1e3 walks vs fp64 oracle (golden)")
    assert count_mismatch_report(batch, f, g["oracle"], "Plummer N=1e3") == 0


@pytest.fixture(scope="module")
def plummer100k():
    batch, _, prm, _ = hz.plummer_case(100000)
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    return batch, prm, ref


def test_plummer100k_default_mode(plummer100k):
    """BASELINE config 2: N = 1e5 equal-mass Plummer, theta 0.3, 200 walks per dispatch."""
    batch, prm, ref = plummer100k
    engine.get_profile(reset=True)
    f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    check_tol(f, ref, "Plummer N=1e5 walks, relative coordinates")
    nbad = count_mismatch_report(batch, f, ref, "Plummer N=1e5")
    assert nbad <= 1e-4 * len(f)
    prof = engine.get_profile()
    I_ep, I_sp = batch.interactions()
    assert prof["n_interaction_ep"] == I_ep and prof["n_interaction_sp"] == I_sp
    assert prof["n_walk"] == batch.n_walk and prof["n_epi"] == batch.n_epi_total
    assert prof["n_call"] == (batch.n_walk + 199) // 200 and prof["n_kernel_launch"] >= 2 * prof["n_call"]
    assert prof["t_calc"] > 0 and prof["t_send"] > 0 and prof["t_recv"] > 0


def test_maximum_walks_per_dispatch(plummer100k):
    """The reference asserts n_walk <= 1000 per dispatch (src/force_gpu_cuda.cu:548); the library has no limit.  1000
    walks per dispatch, and the whole step in one dispatch, with the default 8 sub-batches and the smallest lead."""
    batch, prm, ref = plummer100k
    for k, v in (("streams", 8), ("lead", 3)):
        engine.set_option(k, v)
    f200 = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    for limit in (1000, batch.n_walk):
        engine.get_profile(reset=True)
        f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"], n_walk_limit=limit)
        assert engine.get_profile()["n_call"] == (batch.n_walk + limit - 1) // limit
        check_tol(f, ref, f"n_walk_limit={limit}")
        assert np.array_equal(f["n_ngb"], f200["n_ngb"])
    engine.set_option("lead", 3)


def test_plummer100k_option_matrix(plummer100k):
    """cull on/off and stream count must not change results at all; chunking changes only the
    summation order (<< tolerance); one Newton step must stay within tolerance."""
    batch, prm, ref = plummer100k
    sl = slice(0, 120)
    sub = WalkBatch(batch.epj, batch.spj, batch.epi[:batch.i_off[120]], batch.i_off[:121], batch.id_epj[:batch.ej_off[120]],
                    batch.ej_off[:121], batch.id_spj[:batch.sj_off[120]], batch.sj_off[:121])
    base = engine.calc_force_all_and_write_back(sub, prm["eps"], prm["r_out"], prm["G"])
    r = ref[:sub.n_epi_total]
    engine.set_option("cull", 0)
    f = engine.calc_force_all_and_write_back(sub, prm["eps"], prm["r_out"], prm["G"])
    # cull=0 sends every segment through the exact loop (two-float dx + neighbour test): counts must
    # be identical, forces may differ from the mixed fast/exact default only at fp32 rounding level
    assert np.array_equal(f["n_ngb"], base["n_ngb"]), "neighbour culling changed the counts"
    check_tol(f, r, "cull=0 (exact loop everywhere)")
    assert np.abs(f["acc"] - base["acc"]).max() <= 2e-6 * np.abs(base["acc"]).max()
    engine.set_option("cull", 1)
    for ns in (1, 4):
        engine.set_option("streams", ns)
        f = engine.calc_force_all_and_write_back(sub, prm["eps"], prm["r_out"], prm["G"])
        assert np.array_equal(f["n_ngb"], base["n_ngb"])
        check_tol(f, r, f"streams={ns}")
    engine.set_option("streams", 2)
    engine.set_option("jchunk", 256)
    f = engine.calc_force_all_and_write_back(sub, prm["eps"], prm["r_out"], prm["G"])
    assert np.array_equal(f["n_ngb"], base["n_ngb"])
    check_tol(f, r, "jchunk=256")
    engine.set_option("jchunk", 0)
    engine.set_option("nr", 1)
    f = engine.calc_force_all_and_write_back(sub, prm["eps"], prm["r_out"], prm["G"])
    check_tol(f, r, "one Newton step")
    assert np.array_equal(f["n_ngb"], base["n_ngb"])


def test_absolute_coordinate_mode_matches_reference_kernel_arithmetic(plummer100k):
    """coords=1: dx = float(xj) - float(xi), the arithmetic of reference src/force_gpu_cuda.cu:58-60 that
    src/hard.hpp:1346-1351 replays.  Looser by construction (positions rounded to fp32 before the
    difference): compare against an fp64 evaluation on fp32-rounded positions at full tolerance, and
    against the true oracle at the coordinate-rounding level."""
    batch, prm, ref = plummer100k
    nw = 60
    sub = WalkBatch(batch.epj, batch.spj, batch.epi[:batch.i_off[nw]], batch.i_off[:nw + 1], batch.id_epj[:batch.ej_off[nw]],
                    batch.ej_off[:nw + 1], batch.id_spj[:batch.sj_off[nw]], batch.sj_off[:nw + 1])
    engine.set_option("coords", 1)
    f = engine.calc_force_all_and_write_back(sub, prm["eps"], prm["r_out"], prm["G"])
    engine.set_option("coords", 0)
    rounded = WalkBatch(sub.epj.copy(), sub.spj.copy(), sub.epi.copy(), sub.i_off, sub.id_epj, sub.ej_off, sub.id_spj, sub.sj_off)
    rounded.epj["pos"] = rounded.epj["pos"].astype(np.float32)
    rounded.epi["pos"] = rounded.epi["pos"].astype(np.float32)
    rounded.spj["pos"] = rounded.spj["pos"].astype(np.float32)
    r32 = ob.walks_index(rounded, prm["eps"], prm["r_out"], prm["G"])
    check_tol(f, r32, "absolute fp32 coordinates vs fp64 on fp32-rounded positions")
    ea, _ = rel_err(f, ref[:sub.n_epi_total])
    print(f"[abs coords vs true oracle] acc median {np.median(ea):.3e} max {ea.max():.3e}")
    assert np.median(ea) < 2e-5


def test_direct_mode_equals_index_mode():
    """Non-index functor CalcForceWithLinearCutoffCUDA (per-walk j arrays), reference :704-827."""
    batch, _, prm, _ = hz.plummer_case(1000)
    f_idx = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    f = np.zeros(batch.n_epi_total, dtype=ForceSoft)
    t = batch.pointer_tables(f)
    ej = [np.ascontiguousarray(batch.epj[batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]]]) for w in range(batch.n_walk)]
    sj = [np.ascontiguousarray(batch.spj[batch.id_spj[batch.sj_off[w]:batch.sj_off[w + 1]]]) for w in range(batch.n_walk)]
    ejp = np.array([a.ctypes.data for a in ej], dtype=np.uint64)
    sjp = np.array([a.ctypes.data for a in sj], dtype=np.uint64)
    disp = engine.CalcForceWithLinearCutoffCUDA(0, prm["eps"] ** 2, prm["r_out"] ** 2, prm["G"])
    assert disp(0, t.n_walk, t.epi_ptrs, t.n_epi, ejp, t.n_epj, sjp, t.n_spj) == 0
    assert engine.RetrieveForceCUDA(0, t.n_walk, t.n_epi, t.force_ptrs, direct=True) == 0
    assert np.array_equal(f["n_ngb"], f_idx["n_ngb"])
    assert np.allclose(f["acc"], f_idx["acc"], rtol=1e-12, atol=0) and np.allclose(f["pot"], f_idx["pot"], rtol=1e-12)


def test_ragged_and_tiny_walks():
    """1-particle groups, group sizes that are not multiples of 32, lists shorter than a tile, lists that
    straddle tile boundaries, zero-mass j and type-0 i (GPU/NoSimd semantics: computed like any other)."""
    rng = np.random.default_rng(5)
    n = 3000
    pos = rng.normal(size=(n, 3))
    mass = rng.random(n) / n
    mass[rng.random(n) < 0.1] = 0.0
    rs = 0.02 + 0.05 * rng.random(n)
    epj = _epj(pos, mass, rs)
    spj = np.zeros(500, dtype=SPJQuad)
    spj["mass"] = rng.random(500) / 500
    spj["pos"] = rng.normal(size=(500, 3)) * 0.3 + 8.0
    spj["quad"] = rng.normal(size=(500, 6)) * 1e-3
    sizes_i = [1, 2, 31, 32, 33, 63, 65, 100, 255, 256, 257, 300, 511, 512, 7]
    sizes_e = [1, 7, 8, 9, 255, 256, 257, 511, 513, 1000, 2049, 3000, 0, 5, 64]
    sizes_s = [0, 1, 2, 3, 127, 128, 129, 255, 257, 500, 0, 33, 500, 1, 64]
    i_off, ej_off, sj_off, ide, ids, epi_idx = [0], [0], [0], [], [], []
    for ni, ne, ns in zip(sizes_i, sizes_e, sizes_s):
        epi_idx.append(rng.choice(n, ni, replace=False))
        ide.append(np.sort(rng.choice(n, ne, replace=False)))
        ids.append(rng.choice(500, ns, replace=False))
        i_off.append(i_off[-1] + ni); ej_off.append(ej_off[-1] + ne); sj_off.append(sj_off[-1] + ns)
    epi_idx = np.concatenate(epi_idx)
    epi = _epi(pos[epi_idx], rs[epi_idx])
    epi["type"][::5] = 0
    batch = WalkBatch(epj, spj, epi, i_off, np.concatenate(ide), ej_off, np.concatenate(ids), sj_off)
    ref = ob.walks_index(batch, 1e-3, 0.03, 0.7)
    f = engine.calc_force_all_and_write_back(batch, 1e-3, 0.03, 0.7, n_walk_limit=4)
    nz = np.linalg.norm(ref["acc"], axis=1) > 0
    ea = np.linalg.norm(f["acc"] - ref["acc"], axis=1)[nz] / np.linalg.norm(ref["acc"], axis=1)[nz]
    assert np.median(ea) < 3e-6 and ea.max() < TOL_MAX
    assert np.array_equal(f[~nz]["acc"], ref[~nz]["acc"])
    assert count_mismatch_report(batch, f, ref, "ragged") <= 2
    # the C++ loop driver in the shim and the call-by-call Python loop drive the same functors
    f_py = engine.calc_force_all_and_write_back(batch, 1e-3, 0.03, 0.7, n_walk_limit=4, python_loop=True)
    assert f_py.tobytes() == f.tobytes()


def test_config1_standin_star_cluster():
    """BASELINE configs[0] (sample/star_cluster.sh: N = 1000, Kroupa IMF, 95 % binaries) stand-in: the soft tree sees
    up to 1000 + 12 * 475 = 6700 particles, most groups dominated by zero-mass artificial particles."""
    batch, _, prm, P = hz.kroupa_binary_case(1000, f_bin=0.95)
    assert P["n_bin"] == 475 and len(P["mass"]) == 1000 + 12 * 475
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    check_tol(f, ref, "config 1 stand-in")
    assert count_mismatch_report(batch, f, ref, "config 1 stand-in") <= 2
    cells, groups = batch.tree.export_tree()
    f2 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])         # and with device-built lists
    assert np.array_equal(f2["n_ngb"], f["n_ngb"]) and np.abs(f2["acc"] - f["acc"]).max() <= 2e-6 * np.abs(f["acc"]).max()


def test_one_huge_walk_and_every_ragged_block_size():
    """One walk far beyond PeTar's n_group_limit (2069 i-particles: 64 full blocks and a ragged one of 21) with long
    lists (many chunks per i-block group), then walks whose last block holds 1..32 particles: every lane-sharing
    configuration of the ragged block (4, 2 and 1 lanes per particle) against the oracle."""
    rng = np.random.default_rng(11)
    n = 120000
    pos = rng.normal(scale=0.3, size=(n, 3))
    rs = np.full(n, 2e-3)
    epj = _epj(pos, np.full(n, 1.0 / n), rs)
    spj = np.zeros(40000, dtype=SPJQuad)
    spj["pos"] = rng.normal(scale=3.0, size=(len(spj), 3)) + 6.0
    spj["mass"] = rng.uniform(1e-5, 1e-4, len(spj))
    spj["quad"] = rng.normal(scale=1e-7, size=(len(spj), 6))
    sizes = [2069] + list(range(1, 33)) + [33, 40, 48, 49, 63, 64, 65]
    i_off, ej_off, sj_off, ide, ids, epi_idx = [0], [0], [0], [], [], []
    for k, ni in enumerate(sizes):
        c = rng.integers(0, n)
        near = np.argsort(((pos - pos[c]) ** 2).sum(1))[:max(ni, 64)]
        epi_idx.append(near[:ni])
        ne, ns = (100000, 40000) if k == 0 else (int(rng.integers(300, 3000)), int(rng.integers(0, 2000)))
        ide.append(np.union1d(near, rng.choice(n, ne, replace=False)))            # own neighbourhood + random far j
        ids.append(rng.choice(len(spj), ns, replace=False))
        i_off.append(i_off[-1] + ni); ej_off.append(ej_off[-1] + len(ide[-1])); sj_off.append(sj_off[-1] + ns)
    epi_idx = np.concatenate(epi_idx)
    epi = _epi(pos[epi_idx], rs[epi_idx])
    batch = WalkBatch(epj, spj, epi, i_off, np.concatenate(ide), ej_off, np.concatenate(ids), sj_off)
    ref = ob.walks_index(batch, 0.0, 5e-3, 1.0)
    for streams in (1, 8):
        engine.set_option("streams", streams)
        f = engine.calc_force_all_and_write_back(batch, 0.0, 5e-3, 1.0)
        ea = np.linalg.norm(f["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
        assert np.median(ea) < 1e-6 and ea.max() < TOL_MAX, (streams, np.median(ea), ea.max())
        assert count_mismatch_report(batch, f, ref, f"huge+ragged streams={streams}") <= 2


def test_protocol_errors():
    L = engine.load()
    epi, epj, spj = ob.simdtest_inputs(64, 128, 16)
    batch = WalkBatch.single(epi, epj, spj)
    f = np.zeros(64, dtype=ForceSoft)
    t = batch.pointer_tables(f)
    engine.calc_force_all_and_write_back(batch, 0.0, 0.01, 1.0)       # leaves nothing outstanding
    assert L.pb_retrieve(1, t.n_epi.ctypes.data, t.force_ptrs.ctypes.data, C.byref(engine.LAYOUT_FORCE)) == -4
    rc = L.pb_dispatch_index(1, t.epi_ptrs.ctypes.data, t.n_epi.ctypes.data, C.byref(engine.LAYOUT_EPI), t.id_epj_ptrs.ctypes.data,
                             t.n_epj.ctypes.data, t.id_spj_ptrs.ctypes.data, t.n_spj.ctypes.data)
    assert rc == 0
    rc2 = L.pb_dispatch_index(1, t.epi_ptrs.ctypes.data, t.n_epi.ctypes.data, C.byref(engine.LAYOUT_EPI), t.id_epj_ptrs.ctypes.data,
                              t.n_epj.ctypes.data, t.id_spj_ptrs.ctypes.data, t.n_spj.ctypes.data)
    assert rc2 == -4 and b"not retrieved" in L.pb_last_error()        # tag_max = 1
    bad_ni = np.array([63], dtype=np.int32)
    assert L.pb_retrieve(1, bad_ni.ctypes.data, t.force_ptrs.ctypes.data, C.byref(engine.LAYOUT_FORCE)) == -3
    assert L.pb_retrieve(1, t.n_epi.ctypes.data, t.force_ptrs.ctypes.data, C.byref(engine.LAYOUT_FORCE)) == 0


def test_replay_is_deterministic_and_counts_launches(plummer100k):
    batch, prm, ref = plummer100k
    L = engine.load()
    engine.check(L.pb_record_begin(), "record_begin")
    a = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    engine.check(L.pb_record_end(), "record_end")
    b = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    assert np.array_equal(a, b), "results are not bitwise reproducible run to run"
    ms_t, ms_f = C.c_float(0), C.c_float(0)
    engine.check(L.pb_replay(3, C.byref(ms_t), C.byref(ms_f)), "replay")
    assert 0 < ms_f.value <= ms_t.value * 1.05
    assert L.pb_replay_launches() == 2 * 2 * ((batch.n_walk + 199) // 200)    # 2 kernels x 2 streams x dispatches
    I_ep, I_sp = batch.interactions()
    print(f"[replay N=1e5] {ms_t.value:.3f} ms per step, force kernels {ms_f.value:.3f} ms -> {(I_ep + I_sp) / ms_f.value * 1e-6:.1f} Gint/s")


def test_kroupa_binaries_with_artificial_particles():
    """BASELINE config 3 stand-in at N = 2e4 stars: Kroupa masses (per-particle r_out / r_search),
    10 % binaries each with 11 zero-mass and 3 massive type-0 artificial particles in a tight clump.
    GPU semantics = NoSimd oracle semantics: zero-mass j are counted, type-0 i get a force."""
    batch, _, prm, P = hz.kroupa_binary_case(20000)
    assert (P["mass"] == 0).sum() == 11 * P["n_bin"] and (P["ptype"] == 0).sum() == 3 * P["n_bin"]
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    check_tol(f, ref, "Kroupa + binaries + artificial particles, N=2e4 stars")
    nbad = count_mismatch_report(batch, f, ref, "Kroupa + binaries")
    assert nbad <= 1e-4 * len(f)
    assert f["n_ngb"].max() >= 14          # members + artificial particles of a binary see each other


def test_config4_all_binaries():
    """BASELINE configs[3] / [4] shape (100 % primordial binaries): every star is a binary member, so the soft tree sees 7
    particles per star — per binary 2 zero-mass members, 8 zero-mass probes, 1 zero-mass c.m. and 3 massive type-0
    orbit samples (N = 2e4 stars -> 1.4e5 tree particles, none of them an ordinary single).  Kernel level against the
    NoSimd oracle (coords = 0, this module's fixture), then the library default (coords = 2) on the corrected force,
    through the functor path and through the device-resident tree step."""
    from oracle.dropin_check import DropinChecker
    batch, epi_src, prm, P = hz.kroupa_binary_case(20000, f_bin=1.0)
    n_bin = P["n_bin"]
    assert n_bin == 10000 and len(P["mass"]) == 14 * n_bin and (P["mass"] > 0).sum() == 3 * n_bin
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    check_tol(f, ref, "config 4 shape: 100 % binaries, N=2e4 stars")
    assert count_mismatch_report(batch, f, ref, "100 % binaries") <= 1e-4 * len(f)
    assert f["n_ngb"].min() >= 14                     # everybody sits in a clump of 14
    chk = DropinChecker(P, prm, subset=epi_src)
    cells, groups = batch.tree.export_tree()
    engine.set_option("coords", 2)
    try:
        f2 = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"]).copy()
        f3 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], resident=True).copy()
    finally:
        engine.set_option("coords", 0)
    for name, g in (("functors", f2), ("resident tree step", f3)):
        rep = chk.compare(g, ref)
        print(f"[config 4 shape, drop-in, {name}] {rep}")
        assert rep["acc_rel_err"]["median"] <= 1e-6 and rep["acc_rel_err"]["p99"] <= 1e-5 and rep["acc_rel_err"]["max"] <= 1e-3
        assert rep["pot_tot_rel_err"]["median"] <= 1e-6 and rep["pot_tot_rel_err"]["max"] <= 1e-4
        assert np.array_equal(g["n_ngb"], ref["n_ngb"])
