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
