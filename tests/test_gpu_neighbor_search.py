"""neighbour search for PeTar's second tree (tree_nb) on the GPU, through the
extension functor SearchNeighborCUDAMultiWalk + RetrieveForceCUDA, against the fp64 oracle
SearchNeighborEpEpNoSimd (reference src/soft_force.hpp:11-34).  Integer work: counts must be
bit-exact except for pairs within fp32 rounding of the search radius (counted and reported)."""
import time

import numpy as np
import pytest

from petar_b200 import engine, harness as hz
from petar_b200.types import ForceSoft
from oracle import binding as ob

pytestmark = pytest.mark.gpu


def _oracle_counts(batch):
    """SearchNeighborEpEpNoSimd over each walk's EP list (what FDPS would hand tree_nb's functor)."""
    out = np.zeros(batch.n_epi_total, dtype=np.int64)
    for w in range(batch.n_walk):
        i0, i1 = batch.i_off[w], batch.i_off[w + 1]
        e = batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]]
        out[i0:i1] = ob.search_neighbor(np.ascontiguousarray(batch.epi[i0:i1]), np.ascontiguousarray(batch.epj[e]))["n_ngb"]
    return out


@pytest.mark.parametrize("case", ["plummer", "kroupa_binaries"])
def test_neighbor_counts_vs_oracle(case):
    batch, _, prm, _ = hz.plummer_case(30000) if case == "plummer" else hz.kroupa_binary_case(20000)
    ref = _oracle_counts(batch)
    f = np.zeros(batch.n_epi_total, dtype=ForceSoft)
    f["acc"], f["pot"] = 1.5, -2.5                       # must be left untouched by the count-only retrieve
    engine.tree_neighbor_search(batch, force=f)
    bad = np.nonzero(f["n_ngb"] != ref)[0]
    print(f"[tree_nb {case}] n = {len(ref)}, mean count {ref.mean():.2f}, mismatches {len(bad)}")
    assert len(bad) <= 1e-4 * len(ref)
    for k in bad:                                        # every mismatch must be a borderline pair
        w = np.searchsorted(batch.i_off, k, side="right") - 1
        e = batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]]
        d = batch.epj["pos"][e] - batch.epi["pos"][k]
        r2 = (d * d).sum(1)
        rs2 = np.maximum(batch.epj["r_search"][e], batch.epi["r_search"][k]) ** 2
        assert (np.abs(r2 - rs2) <= 4e-6 * rs2).sum() >= abs(int(f["n_ngb"][k]) - int(ref[k]))
    assert np.all(f["acc"] == 1.5) and np.all(f["pot"] == -2.5)
    # and it agrees with the count the force kernel produces on the same lists (eps = 0)
    ff = engine.calc_force_all_and_write_back(batch, 0.0, prm["r_out"], prm["G"])
    assert np.array_equal(ff["n_ngb"], f["n_ngb"])


def test_neighbor_search_throughput_report():
    batch, _, prm, _ = hz.plummer_case(300000)
    engine.tree_neighbor_search(batch)                   # warm-up
    t0 = time.perf_counter()
    f = engine.tree_neighbor_search(batch)
    dt = time.perf_counter() - t0
    pairs = batch.interactions()[0]
    t1 = time.perf_counter()
    nw = 64
    sub = _oracle_counts(type(batch)(batch.epj, batch.spj, batch.epi[:batch.i_off[nw]], batch.i_off[:nw + 1], batch.id_epj[:batch.ej_off[nw]],
                                     batch.ej_off[:nw + 1], batch.id_spj[:0], np.zeros(nw + 1, dtype=np.int64)))
    dt_cpu = time.perf_counter() - t1
    assert np.array_equal(sub, f["n_ngb"][:len(sub)])
    print(f"[tree_nb N=3e5] {pairs:.3e} candidate pairs in {dt * 1e3:.1f} ms end to end = {pairs / dt * 1e-9:.0f} Gpairs/s "
          f"(fp64 scalar oracle on the first {nw} walks: {dt_cpu:.2f} s)")


def _true_pairs(batch):
    """All (i, j) with |x_i - x_j| < max(rs_i, rs_j) in fp64 (i over batch.epi, j over batch.epj), as a sorted
    array of keys i * n_j + j, plus the candidates' relative distance to the search radius."""
    from scipy.spatial import cKDTree
    pi, pj = batch.epi["pos"], batch.epj["pos"]
    ri, rj = batch.epi["r_search"], batch.epj["r_search"]
    nj = len(pj)
    a = cKDTree(pj).query_ball_point(pi, ri * (1 + 1e-9), workers=-1)            # j within rs_i of i
    b = cKDTree(pi).query_ball_point(pj, rj * (1 + 1e-9), workers=-1)            # i within rs_j of j
    ii = np.concatenate([np.repeat(np.arange(len(pi)), [len(h) for h in a]), np.fromiter((i for h in b for i in h), dtype=np.int64)])
    jj = np.concatenate([np.fromiter((j for h in a for j in h), dtype=np.int64), np.repeat(np.arange(nj), [len(h) for h in b])])
    key = np.unique(ii * nj + jj)
    ii, jj = key // nj, key % nj
    d2 = ((pi[ii] - pj[jj]) ** 2).sum(1)
    rs2 = np.maximum(ri[ii], rj[jj]) ** 2
    return key[d2 < rs2], nj


@pytest.mark.parametrize("case", ["plummer", "kroupa_binaries"])
def test_neighbor_lists_vs_fp64_search(case):
    """Option nb_lists: the pairs behind the counts, i.e. what getNeighborListOneParticle returns particle by
    particle.  Index work: the lists must be the fp64 sets except for pairs within fp32 rounding of the radius."""
    batch, _, prm, _ = hz.plummer_case(30000) if case == "plummer" else hz.kroupa_binary_case(20000, f_bin=0.2)
    f, off, idx = engine.tree_neighbor_search(batch, lists=True)
    assert off[0] == 0 and off[-1] == len(idx) and np.array_equal(np.diff(off), f["n_ngb"])
    assert engine.tree_neighbor_search(batch)["n_ngb"].tobytes() == f["n_ngb"].tobytes()      # same counts without lists
    want, nj = _true_pairs(batch)
    got = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off)) * nj + idx
    assert np.all(np.diff(got) > 0)                                                           # ascending i, then ascending j, no duplicates
    odd = np.setxor1d(got, want)
    print(f"[nb lists {case}] {len(got)} pairs, {len(odd)} differ from the fp64 sets")
    assert len(odd) <= 1e-5 * len(want) + 2
    for key in odd:                                                                           # borderline pairs only
        i, j = key // nj, key % nj
        d2 = ((batch.epi["pos"][i] - batch.epj["pos"][j]) ** 2).sum()
        rs2 = max(batch.epi["r_search"][i], batch.epj["r_search"][j]) ** 2
        assert abs(d2 - rs2) <= 4e-6 * rs2
    assert np.diff(off).min() >= 1                                                            # every particle finds itself
    # the lists are what the correction consumes: run it on them against the oracle
    from petar_b200.types import PtclCorr
    pj = np.zeros(len(batch.epj), dtype=PtclCorr)
    for k in ("id", "mass", "pos", "r_in", "r_out"):
        pj[k] = batch.epj[k]
    pi = np.zeros(batch.n_epi_total, dtype=PtclCorr)
    pi["id"], pi["pos"] = batch.epi["id"], batch.epi["pos"]
    src = {int(v): k for k, v in enumerate(batch.epj["id"])}
    at = np.array([src[int(v)] for v in batch.epi["id"]])
    for k in ("mass", "r_in", "r_out"):
        pi[k] = batch.epj[k][at]
    ref = ob.correct_force_tree_neighbor(pi.copy(), off, idx, pj, 0.0, prm["r_out"], 1.0, False)
    out = engine.correct_force_with_cutoff_tree_neighbor(pi.copy(), off, idx, pj, 0.0, prm["r_out"], 1.0, False)
    assert out.tobytes() == ref.tobytes() and not np.array_equal(out["acc"], pi["acc"])


def test_neighbor_lists_overflow_regrows():
    """A dense clump: far more pairs than the initial buffer estimate (12 per particle) — the sub-batch is re-run."""
    rng = np.random.default_rng(2)
    n = 4000
    pos = rng.normal(scale=1e-3, size=(n, 3))
    rs = np.full(n, 2e-3)
    batch, _ = hz.build_walk_batch(pos, np.full(n, 1.0 / n), rs)
    f, off, idx = engine.tree_neighbor_search(batch, lists=True)
    assert np.array_equal(np.diff(off), f["n_ngb"]) and f["n_ngb"].mean() > 200
    want, nj = _true_pairs(batch)
    got = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off)) * nj + idx
    assert len(np.setxor1d(got, want)) <= 1e-4 * len(want)
