"""Full-size checks at BASELINE.json's N = 1e6 (config 3 stand-in: Kroupa masses, 10 % binaries with
artificial particles, 1.6e6 tree particles, ~2.3e10 interactions per tree step), through size-independent
properties of the force law plus an oracle comparison on a sample of walks:

* scaling all masses by 4 scales acc and pot by exactly 4 (powers of two commute with every fp32
  rounding in the kernels) and leaves the neighbour counts unchanged — bitwise;
* translating every position by a constant vector leaves acc, pot and counts unchanged up to the
  rounding of the fp64 inputs (this is what the walk-origin-relative hi/lo coordinates are for);
* two runs are bitwise identical (fixed summation order, no atomics);
* 96 walks sampled across the step agree with the fp64 oracle within the tolerance."""
import numpy as np
import pytest

from petar_b200 import engine, harness as hz
from petar_b200.walks import WalkBatch
from oracle import binding as ob

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("coords0")]   # kernel-level parity: see conftest.coords0


_FULL = {}


@pytest.fixture(scope="module")
def full():
    batch, epi_src, prm, P = hz.kroupa_binary_case(1000000)
    engine.set_option("coords", 2)                       # the library default: the drop-in mode, checked on the corrected force
    f_dropin = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"]).copy()
    engine.set_option("coords", 0)                       # kernel-level parity statements below
    f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    _FULL.update(epi_src=epi_src, P=P, f_dropin=f_dropin)
    return batch, prm, f


def test_fullsize_shape(full):
    batch, prm, f = full
    I_ep, I_sp = batch.interactions()
    print(f"[N=1e6 config 3 stand-in] tree particles {len(batch.epj)}, walks {batch.n_walk}, EP-EP {I_ep:.3e}, EP-SP {I_sp:.3e}")
    assert len(batch.epj) == 1600000 and batch.n_epi_total == 1600000
    assert np.isfinite(f["acc"]).all() and np.isfinite(f["pot"]).all()
    assert f["n_ngb"].min() >= 1                      # everybody counts itself


def test_fullsize_oracle_sample(full):
    batch, prm, f = full
    rng = np.random.default_rng(0)
    ws = np.sort(rng.choice(batch.n_walk, 96, replace=False))
    ea_all, ep_all, nbad = [], [], 0
    for w in ws:
        i0, i1 = batch.i_off[w], batch.i_off[w + 1]
        sub = WalkBatch(batch.epj, batch.spj, batch.epi[i0:i1], [0, i1 - i0], batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]],
                        [0, batch.ej_off[w + 1] - batch.ej_off[w]], batch.id_spj[batch.sj_off[w]:batch.sj_off[w + 1]],
                        [0, batch.sj_off[w + 1] - batch.sj_off[w]])
        ref = ob.walks_index(sub, prm["eps"], prm["r_out"], prm["G"])
        g = f[i0:i1]
        ea_all.append(np.linalg.norm(g["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1))
        ep_all.append(np.abs((g["pot"] - ref["pot"]) / ref["pot"]))
        nbad += int((g["n_ngb"] != ref["n_ngb"]).sum())
    ea, ep = np.concatenate(ea_all), np.concatenate(ep_all)
    print(f"[N=1e6, 96 sampled walks, {len(ea)} particles] acc rel err median {np.median(ea):.3e} p99 {np.percentile(ea, 99):.3e} "
          f"max {ea.max():.3e} | pot median {np.median(ep):.3e} max {ep.max():.3e} | n_ngb mismatches {nbad}")
    assert np.median(ea) <= 1e-6 and ea.max() <= 1e-4
    assert np.median(ep) <= 1e-6 and ep.max() <= 1e-4
    assert nbad <= 1e-4 * len(ea)


def test_fullsize_dropin_mode_sample(full):
    """The library default (coords = 2) at full size: kernel force + the float-replay correction an unmodified PeTar
    applies against fp64 oracle + all-double correction, on the particles of 96 sampled walks (oracle/dropin_check.py)."""
    from oracle.dropin_check import DropinChecker
    batch, prm, f0 = full
    fd, epi_src, P = _FULL["f_dropin"], _FULL["epi_src"], _FULL["P"]
    rng = np.random.default_rng(0)
    ws = np.sort(rng.choice(batch.n_walk, 96, replace=False))
    rows = np.concatenate([np.arange(batch.i_off[w], batch.i_off[w + 1]) for w in ws])
    refs = []
    for w in ws:
        i0, i1 = batch.i_off[w], batch.i_off[w + 1]
        sub = WalkBatch(batch.epj, batch.spj, batch.epi[i0:i1], [0, i1 - i0], batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]],
                        [0, batch.ej_off[w + 1] - batch.ej_off[w]], batch.id_spj[batch.sj_off[w]:batch.sj_off[w + 1]],
                        [0, batch.sj_off[w + 1] - batch.sj_off[w]])
        refs.append(ob.walks_index(sub, prm["eps"], prm["r_out"], prm["G"]))
    ref = np.concatenate(refs)
    chk = DropinChecker(P, prm, subset=epi_src[rows])
    rep = chk.compare(fd[rows], ref)
    rep0 = chk.compare(f0[rows], ref)                    # the walk-relative kernel with the same unmodified replay, for contrast
    print(f"[N=1e6 drop-in, 96 walks] coords=2 + float replay: {rep}")
    print(f"[N=1e6 drop-in, 96 walks] coords=0 + float replay: acc {rep0['acc_rel_err']} residual/neighbour {rep0['abs_residual_per_neighbour']}")
    assert np.array_equal(fd["n_ngb"], f0["n_ngb"])
    assert rep["pass"], rep
    assert rep["abs_residual_per_neighbour"]["median"] < 0.2 * rep0["abs_residual_per_neighbour"]["median"]


def test_fullsize_determinism_and_mass_scaling(full):
    batch, prm, f = full
    f2 = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    assert np.array_equal(f, f2), "two runs differ bitwise"
    heavy = WalkBatch(batch.epj.copy(), batch.spj.copy(), batch.epi, batch.i_off, batch.id_epj, batch.ej_off, batch.id_spj, batch.sj_off)
    heavy.epj["mass"] *= 4.0
    heavy.spj["mass"] *= 4.0
    heavy.spj["quad"] *= 4.0
    f4 = engine.calc_force_all_and_write_back(heavy, prm["eps"], prm["r_out"], prm["G"])
    assert np.array_equal(f4["n_ngb"], f["n_ngb"])
    assert np.array_equal(f4["acc"], 4.0 * f["acc"]) and np.array_equal(f4["pot"], 4.0 * f["pot"])


def test_fullsize_translation_invariance(full):
    batch, prm, f = full
    shift = np.array([1.0, -2.0, 0.5])
    moved = WalkBatch(batch.epj.copy(), batch.spj.copy(), batch.epi.copy(), batch.i_off, batch.id_epj, batch.ej_off, batch.id_spj, batch.sj_off)
    moved.epj["pos"] += shift
    moved.spj["pos"] += shift
    moved.epi["pos"] += shift
    fm = engine.calc_force_all_and_write_back(moved, prm["eps"], prm["r_out"], prm["G"])
    ea = np.linalg.norm(fm["acc"] - f["acc"], axis=1) / np.linalg.norm(f["acc"], axis=1)
    ep = np.abs((fm["pot"] - f["pot"]) / f["pot"])
    nbad = int((fm["n_ngb"] != f["n_ngb"]).sum())
    print(f"[N=1e6 translated by {shift.tolist()}] acc rel change median {np.median(ea):.3e} max {ea.max():.3e} | pot max {ep.max():.3e} | n_ngb changes {nbad}")
    assert np.median(ea) <= 1e-6 and ea.max() <= 1e-4 and ep.max() <= 1e-4
    assert nbad <= 1e-5 * len(f)
