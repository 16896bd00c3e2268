"""GPU parity of the changeover correction: pb_correct_changeover against the CPU oracle
(oracle/oracle_changeover.c, itself pinned bit for bit against the reference's function in tests/test_oracle.py).
The bar is bit-exactness: the kernel is fp64 (fp32 for the float replay), compiled without FMA contraction,
and applies the pairs in list order."""
import numpy as np
import pytest

from petar_b200 import engine, harness
from petar_b200.types import PtclCorr, LARGE_FLOAT
from petar_b200.walks import WalkBatch
from oracle import binding as ob

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("coords0")]   # kernel-level parity: see conftest.coords0


def _random_set(n, r_out_g, seed, clump=0.3):
    rng = np.random.default_rng(seed)
    p = np.zeros(n, dtype=PtclCorr)
    p["id"] = rng.permutation(n) + 1
    # clumps of a few particles well inside the changeover region, plus a uniform background
    centres = rng.uniform(-1, 1, (max(1, n // 4), 3))
    p["pos"] = centres[rng.integers(0, len(centres), n)] + rng.normal(scale=clump * r_out_g, size=(n, 3))
    p["mass"] = 10 ** rng.uniform(-7, -4, n)
    f = 10 ** rng.uniform(0, 0.5, n)
    p["r_in"], p["r_out"] = 0.1 * r_out_g * f, r_out_g * f
    kind = rng.integers(0, 5, n)
    mem = kind == 1
    p["status"][mem] = -(rng.integers(1, n, mem.sum()) + 0.0)
    p["mass_backup"][mem] = p["mass"][mem]
    p["mass"][mem] = 0.0
    orphan = kind == 2                                        # member without c.m. particle
    p["status"][orphan] = -LARGE_FLOAT
    p["mass_backup"][orphan] = p["mass"][orphan]
    art = kind == 3
    p["status"][art] = rng.integers(1, 9, art.sum()) + 0.0
    p["acc"] = rng.normal(size=(n, 3))
    p["pot_tot"], p["pot_soft"] = -rng.uniform(0.5, 2, n), -rng.uniform(0.5, 2, n)
    return p


@pytest.mark.parametrize("replay_fp32", [False, True])
@pytest.mark.parametrize("eps", [0.0, 1e-4])
def test_random_pairs_bit_exact(replay_fp32, eps):
    r_out_g = 2e-3
    p = _random_set(6000, r_out_g, 3)
    off, idx = harness.neighbor_lists(p["pos"], 3.0 * p["r_out"])
    assert np.diff(off).max() >= 4 and (np.diff(off) == 1).any()      # clumps and isolated particles (self only)
    ref = ob.correct_force_tree_neighbor(p.copy(), off, idx, p, eps, r_out_g, 0.7, replay_fp32)
    got = engine.correct_force_with_cutoff_tree_neighbor(p.copy(), off, idx, p, eps, r_out_g, 0.7, replay_fp32)
    assert got.tobytes() == ref.tobytes()
    assert not np.array_equal(got["acc"], p["acc"])


def test_kroupa_binaries_bit_exact_and_idempotent_inputs():
    P = harness.kroupa_binary_particles(20000, f_bin=0.2)
    prm = P["prm"]
    p = harness.corr_particles(P)
    off, idx = harness.neighbor_lists(P["pos"], P["rs"])
    for replay in (False, True):
        ref = ob.correct_force_tree_neighbor(p.copy(), off, idx, p, 0.0, prm["r_out"], 1.0, replay)
        got = engine.correct_force_with_cutoff_tree_neighbor(p.copy(), off, idx, p, 0.0, prm["r_out"], 1.0, replay)
        assert got.tobytes() == ref.tobytes()
    # inputs other than acc / pot are untouched, and a second array layout (EPJSoft-like neighbours) binds too
    assert all(np.array_equal(got[k], p[k]) for k in ("id", "mass", "pos", "r_in", "r_out", "mass_backup", "status"))


def test_empty_and_errors():
    p = _random_set(10, 2e-3, 1)
    off = np.zeros(11, dtype=np.int32)
    out = engine.correct_force_with_cutoff_tree_neighbor(p.copy(), off, np.zeros(0, np.int32), p, 0.0, 2e-3, 1.0)
    ref = ob.correct_force_tree_neighbor(p.copy(), off, np.zeros(0, np.int32), p, 0.0, 2e-3, 1.0, False)
    assert out.tobytes() == ref.tobytes()                     # self-potential term only
    engine.correct_force_with_cutoff_tree_neighbor(p[:0].copy(), np.zeros(1, np.int32), np.zeros(0, np.int32), p, 0.0, 2e-3, 1.0)
    bad = np.arange(11, dtype=np.int32)
    with pytest.raises(engine.PbError):
        engine.correct_force_with_cutoff_tree_neighbor(p.copy(), bad, np.full(10, 99, np.int32), p, 0.0, 2e-3, 1.0)


def test_soft_force_plus_correction_is_the_changeover_force():
    """Linear-cutoff force from the force kernel + the correction = the changeover-weighted Newtonian sum
    (k = 1 - W0 of the pair member with the larger r_out; k = 1 beyond r_out), evaluated here in fp64."""
    n, G = 1500, 1.0
    rng = np.random.default_rng(5)
    P = harness.kroupa_binary_particles(n, f_bin=0.0)
    prm = P["prm"]
    p = harness.corr_particles(P)
    # squeeze the system so that many pairs sit inside the changeover region
    p["pos"] *= 20 * prm["r_out"]
    rs = np.maximum(P["rs"], 1.5 * p["r_out"])
    from petar_b200.types import EPISoft, EPJSoft, SPJQuad
    epi = np.zeros(n, dtype=EPISoft); epj = np.zeros(n, dtype=EPJSoft)
    epi["id"] = epj["id"] = p["id"]; epi["pos"] = epj["pos"] = p["pos"]; epi["r_search"] = epj["r_search"] = rs
    epi["type"] = 1; epj["mass"] = p["mass"]; epj["r_in"], epj["r_out"] = p["r_in"], p["r_out"]
    f = engine.calc_force_all_and_write_back(WalkBatch.single(epi, epj, np.zeros(0, dtype=SPJQuad)), 0.0, prm["r_out"], G)
    p["acc"] = f["acc"]
    off, idx = harness.neighbor_lists(p["pos"], rs)
    got = engine.correct_force_with_cutoff_tree_neighbor(p.copy(), off, idx, p, 0.0, prm["r_out"], G)
    # fp64 direct evaluation
    want = np.zeros((n, 3))
    for i in range(n):
        dr = p["pos"][i] - p["pos"]
        r = np.sqrt((dr ** 2).sum(axis=1)); r[i] = 1.0
        big = p["r_out"] > p["r_out"][i]
        r_in = np.where(big, p["r_in"], p["r_in"][i]); r_o = np.where(big, p["r_out"], p["r_out"][i])
        k = np.array([1.0 - ob.changeover_w(a, b, c)[0] for a, b, c in zip(r_in, r_o, r)])
        w = G * p["mass"] * k / r ** 3; w[i] = 0.0
        want[i] = -(w[:, None] * dr).sum(axis=0)
    err = np.linalg.norm(got["acc"] - want, axis=1) / np.linalg.norm(want, axis=1)
    assert np.median(err) < 1e-6 and err.max() < 1e-4, (np.median(err), err.max())


def test_kernel_vs_committed_reference_vectors():
    """The device kernel against outputs of the REFERENCE's own pair function (tests/golden/changeover_pairs.npz,
    generated where /root/reference exists): one neighbour per particle, both branches, eps = 0 and 1e-4."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "changeover_pairs.npz"))
    pi, pj, eps = g["pi"], g["pj"], g["eps"]
    for replay, key in ((False, "ref_fp64"), (True, "ref_replay_fp32")):
        for e in np.unique(eps):
            sel = np.nonzero(eps == e)[0]
            off = np.arange(len(sel) + 1, dtype=np.int32)
            got = engine.correct_force_with_cutoff_tree_neighbor(np.ascontiguousarray(pi[sel]), off, sel.astype(np.int32), pj,
                                                                 float(e), float(g["r_out"]), float(g["G"]), replay)
            assert got.tobytes() == np.ascontiguousarray(g[key][sel]).tobytes()
