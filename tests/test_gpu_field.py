"""the direct-sum field query behind AMUSE's get_gravity_at_point /
get_potential_at_point (reference amuse-interface/interface.cc:966-1030 -> CalcForcePPSimd,
src/soft_force.hpp:285-344), on the GPU through pb_field_at_points, against the fp64 oracle
CalcForcePPNoSimd restatement."""
import time

import numpy as np
import pytest

from petar_b200 import engine, harness as hz
from petar_b200.types import EPISoft, EPJSoft
from oracle import binding as ob

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("coords0")]   # kernel-level parity: see conftest.coords0


def _oracle(points, pos, mass, G):
    epi = np.zeros(len(points), dtype=EPISoft)
    epi["pos"] = points
    epj = np.zeros(len(mass), dtype=EPJSoft)
    epj["pos"], epj["mass"] = pos, mass
    return ob.force_pp(epi, epj, G)


def test_field_at_points_vs_oracle():
    mass, pos, _ = hz.make_plummer(50000)
    mass = mass.copy()
    mass[::17] = 0.0                                   # zero-mass particles are skipped (SIMD-path semantics)
    rng = np.random.default_rng(11)
    # query points: inside the cluster, in the halo, far outside, and ragged count (not a multiple of 32/512)
    pts = np.concatenate([rng.normal(size=(700, 3)) * 0.5, rng.normal(size=(300, 3)) * 5.0, rng.normal(size=(37, 3)) * 100.0])
    part = np.zeros(len(mass), dtype=EPJSoft)
    part["pos"], part["mass"] = pos, mass
    G = 0.5
    ax, ay, az, phi = engine.get_gravity_and_potential_at_point(pts[:, 0], pts[:, 1], pts[:, 2], part, G=G)
    ref = _oracle(pts, pos[mass > 0], mass[mass > 0], G)
    acc = np.stack([ax, ay, az], axis=1)
    ea = np.linalg.norm(acc - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ep = np.abs((phi - ref["pot"]) / ref["pot"])
    print(f"[field query, 1037 points x 47k particles] acc rel err median {np.median(ea):.3e} max {ea.max():.3e} | pot median {np.median(ep):.3e} max {ep.max():.3e}")
    assert np.median(ea) <= 1e-6 and ea.max() <= 1e-4
    assert np.median(ep) <= 1e-6 and ep.max() <= 1e-4
    # the soft-force path still works afterwards (the j store is re-published by the next tree step)
    batch, _, prm, _ = hz.plummer_case(1000)
    f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])
    assert np.array_equal(f["n_ngb"], ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])["n_ngb"])


def test_field_query_throughput_report():
    mass, pos, _ = hz.make_plummer(1000000)
    part = np.zeros(len(mass), dtype=EPJSoft)
    part["pos"], part["mass"] = pos, mass
    rng = np.random.default_rng(5)
    pts = rng.normal(size=(20000, 3))
    engine.get_gravity_and_potential_at_point(pts[:64, 0], pts[:64, 1], pts[:64, 2], part)          # warm-up
    t0 = time.perf_counter()
    ax, ay, az, phi = engine.get_gravity_and_potential_at_point(pts[:, 0], pts[:, 1], pts[:, 2], part)
    dt = time.perf_counter() - t0
    sub = slice(0, 256)
    t1 = time.perf_counter()
    ref = _oracle(pts[sub], pos, mass, 1.0)
    dt_cpu = (time.perf_counter() - t1) * len(pts) / 256
    ea = np.linalg.norm(np.stack([ax, ay, az], 1)[sub] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    print(f"[field query 2e4 points x 1e6 particles] GPU end-to-end {dt * 1e3:.1f} ms = {2e4 * 1e6 / dt * 1e-9:.0f} Ginteractions/s "
          f"(fp64 scalar oracle extrapolated: {dt_cpu:.1f} s); acc err median {np.median(ea):.2e} max {ea.max():.2e}")
    assert np.median(ea) <= 1e-6 and ea.max() <= 1e-4


def test_field_query_leaves_the_j_store_unpublished():
    """pb_field_at_points overwrites the j store with the query's particle set; a dispatch that follows without a fresh
    pb_upload_j must fail with PB_ERR_PROTOCOL instead of silently running against it."""
    import ctypes as C
    from petar_b200.types import EPJSoft
    batch, _, prm, _ = hz.plummer_case(2000)
    f0 = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"]).copy()
    part = np.zeros(len(batch.epj), dtype=EPJSoft)
    part["pos"], part["mass"] = batch.epj["pos"], batch.epj["mass"]
    engine.get_gravity_and_potential_at_point(np.zeros(4), np.zeros(4), np.ones(4), part)
    with pytest.raises(AssertionError):            # the shim aborts on PB_ERR_PROTOCOL; the Python loop driver asserts rc == 0
        t = batch.pointer_tables(np.zeros(batch.n_epi_total, dtype=engine.ForceSoft))
        L = engine.load()
        rc = L.pb_dispatch_index(t.n_walk, t.epi_ptrs.ctypes.data, t.n_epi.ctypes.data, C.byref(engine.LAYOUT_EPI),
                                 t.id_epj_ptrs.ctypes.data, t.n_epj.ctypes.data, t.id_spj_ptrs.ctypes.data, t.n_spj.ctypes.data)
        assert rc == 0, L.pb_last_error()
    f1 = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])      # publishes j again
    assert f1.tobytes() == f0.tobytes()
