"""GPU tests of the engine's life cycle and of the layout-agnostic C ABI: re-initialisation,
buffer growth, foreign struct layouts (a KDKDK_4TH-like build of PeTar), the indexed LET packer."""
import ctypes as C

import numpy as np
import pytest

from petar_b200 import engine, harness as hz
from petar_b200.types import ForceSoft
from petar_b200.walks import WalkBatch
from oracle import binding as ob

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("coords0")]   # kernel-level parity: see conftest.coords0


def _tol(f_acc, f_pot, ref):
    ea = np.linalg.norm(f_acc - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ep = np.abs((f_pot - ref["pot"]) / ref["pot"])
    assert np.median(ea) <= 1e-6 and ea.max() <= 1e-4 and np.median(ep) <= 1e-6 and ep.max() <= 1e-4


def test_finalize_and_reinit_and_growth():
    L = engine.load()
    small, _, prm_s, _ = hz.plummer_case(500)
    big, _, prm_b, _ = hz.plummer_case(30000)
    f1 = engine.calc_force_all_and_write_back(small, prm_s["eps"], prm_s["r_out"], prm_s["G"])
    f2 = engine.calc_force_all_and_write_back(big, prm_b["eps"], prm_b["r_out"], prm_b["G"])      # every buffer grows
    L.pb_finalize()
    engine.set_option("streams", 3)
    f3 = engine.calc_force_all_and_write_back(small, prm_s["eps"], prm_s["r_out"], prm_s["G"])    # lazy re-init
    assert np.array_equal(f1["n_ngb"], f3["n_ngb"]) and np.allclose(f1["acc"], f3["acc"], rtol=1e-12, atol=0)
    _tol(f2["acc"], f2["pot"], ob.walks_index(big, prm_b["eps"], prm_b["r_out"], prm_b["G"]))
    engine.set_option("streams", 8)


def test_foreign_struct_layouts():
    """PeTar built with KDKDK_4TH has EPISoft/EPJSoft with an extra `acc` member and ForceSoft with
    `acorr` (reference src/soft_ptcl.hpp:8-10, 279-281, 316-318).  The C ABI takes stride + offsets,
    so such arrays bind without recompiling; members the kernel does not own must stay untouched."""
    L = engine.load()
    batch, _, prm, _ = hz.plummer_case(3000)
    epi4 = np.dtype([("id", "<i8"), ("pos", "<f8", (3,)), ("r_search", "<f8"), ("rank_org", "<i4"), ("type", "<i4"), ("acc", "<f8", (3,))], align=True)
    epj4 = np.dtype([("id", "<i8"), ("mass", "<f8"), ("pos", "<f8", (3,)), ("vel", "<f8", (3,)), ("acc", "<f8", (3,)), ("r_in", "<f8"),
                     ("r_out", "<f8"), ("r_search", "<f8"), ("r_scale_next", "<f8"), ("group_data", "<i8", (2,)), ("rank_org", "<i4"), ("adr_org", "<i4")], align=True)
    frc4 = np.dtype([("acc", "<f8", (3,)), ("pot", "<f8"), ("acorr", "<f8", (3,)), ("n_ngb", "<i8")], align=True)
    ei = np.zeros(len(batch.epi), dtype=epi4)
    ej = np.zeros(len(batch.epj), dtype=epj4)
    for k in ("pos", "r_search"):
        ei[k] = batch.epi[k]
        ej[k] = batch.epj[k]
    ej["mass"] = batch.epj["mass"]
    ei["acc"] = 7.0
    fr = np.zeros(batch.n_epi_total, dtype=frc4)
    fr["acorr"] = 3.25
    lei = engine.LayoutEpi(epi4.itemsize, epi4.fields["pos"][1], epi4.fields["r_search"][1])
    lej = engine.LayoutEpj(epj4.itemsize, epj4.fields["pos"][1], epj4.fields["mass"][1], epj4.fields["r_search"][1])
    lfr = engine.LayoutForce(frc4.itemsize, frc4.fields["acc"][1], frc4.fields["pot"][1], frc4.fields["n_ngb"][1])
    io = batch.i_off[:-1]
    epi_ptrs = (ei.ctypes.data + io * epi4.itemsize).astype(np.uint64)
    frc_ptrs = (fr.ctypes.data + io * frc4.itemsize).astype(np.uint64)
    ide_ptrs = (batch.id_epj.ctypes.data + batch.ej_off[:-1] * 4).astype(np.uint64)
    ids_ptrs = (batch.id_spj.ctypes.data + batch.sj_off[:-1] * 4).astype(np.uint64)
    n_epi, n_epj, n_spj = batch.n_epi, batch.n_epj, batch.n_spj
    engine.check(L.pb_set_params(prm["eps"] ** 2, prm["r_out"] ** 2, prm["G"]), "set_params")
    engine.check(L.pb_upload_j(ej.ctypes.data, len(ej), C.byref(lej), batch.spj.ctypes.data, len(batch.spj), C.byref(engine.LAYOUT_SPJ)), "upload_j")
    engine.check(L.pb_dispatch_index(batch.n_walk, epi_ptrs.ctypes.data, n_epi.ctypes.data, C.byref(lei), ide_ptrs.ctypes.data, n_epj.ctypes.data,
                                     ids_ptrs.ctypes.data, n_spj.ctypes.data), "dispatch_index")
    engine.check(L.pb_retrieve(batch.n_walk, n_epi.ctypes.data, frc_ptrs.ctypes.data, C.byref(lfr)), "retrieve")
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    _tol(fr["acc"], fr["pot"], ref)
    assert np.array_equal(fr["n_ngb"], ref["n_ngb"])
    assert np.all(fr["acorr"] == 3.25), "retrieve wrote outside acc / pot / n_ngb"


def test_monopole_only_superparticles():
    """pb_layout_spj.has_quad = 0 (PS::SPJMonopoleInAndOut, a build without USE_QUAD)."""
    L = engine.load()
    batch, _, prm, _ = hz.plummer_case(2000)
    mono = np.zeros(len(batch.spj), dtype=np.dtype([("mass", "<f8"), ("pos", "<f8", (3,))], align=True))
    mono["mass"], mono["pos"] = batch.spj["mass"], batch.spj["pos"]
    lsp = engine.LayoutSpj(32, 8, 0, 0, 0)
    f = np.zeros(batch.n_epi_total, dtype=ForceSoft)
    t = batch.pointer_tables(f)
    engine.check(L.pb_set_params(prm["eps"] ** 2, prm["r_out"] ** 2, prm["G"]), "set_params")
    engine.check(L.pb_upload_j(batch.epj.ctypes.data, len(batch.epj), C.byref(engine.LAYOUT_EPJ), mono.ctypes.data, len(mono), C.byref(lsp)), "upload_j")
    engine.check(L.pb_dispatch_index(t.n_walk, t.epi_ptrs.ctypes.data, t.n_epi.ctypes.data, C.byref(engine.LAYOUT_EPI), t.id_epj_ptrs.ctypes.data,
                                     t.n_epj.ctypes.data, t.id_spj_ptrs.ctypes.data, t.n_spj.ctypes.data), "dispatch_index")
    engine.check(L.pb_retrieve(t.n_walk, t.n_epi.ctypes.data, t.force_ptrs.ctypes.data, C.byref(engine.LAYOUT_FORCE)), "retrieve")
    zq = WalkBatch(batch.epj, batch.spj.copy(), batch.epi, batch.i_off, batch.id_epj, batch.ej_off, batch.id_spj, batch.sj_off)
    zq.spj["quad"] = 0.0
    ref = ob.walks_index(zq, prm["eps"], prm["r_out"], prm["G"])
    _tol(f["acc"], f["pot"], ref)


def test_raw_upload_equals_host_packing():
    """Option raw_upload: the caller's j arrays are page-locked, copied as they are and packed on the device — the forces
    must be bit-identical to the host-packed path (same fp64 -> hi/lo fp32 arithmetic, no FMA contraction on the device)."""
    batch, _, prm, _ = hz.kroupa_binary_case(6000)
    f0 = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"]).copy()
    engine.set_option("raw_upload", 1)
    try:
        f1 = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"]).copy()
        moved = type(batch)(batch.epj.copy(), batch.spj.copy(), batch.epi, batch.i_off, batch.id_epj, batch.ej_off, batch.id_spj, batch.sj_off)
        f2 = engine.calc_force_all_and_write_back(moved, prm["eps"], prm["r_out"], prm["G"]).copy()      # other host arrays: re-registered
    finally:
        engine.set_option("raw_upload", 0)
    assert f1.tobytes() == f0.tobytes() and f2.tobytes() == f0.tobytes()


def test_ep_lists_as_runs_are_bitwise_equivalent():
    """Option ep_runs (opt-in): the EP index lists cross PCIe as (start, length) runs and are expanded on the device;
    forces, potentials and counts must equal the plain-index path bit for bit — also for lists without any run
    (shuffled j store: every run has length 1), for the neighbour search, and in the device-resident replay."""
    batch, _, prm, _ = hz.kroupa_binary_case(6000)
    rng = np.random.default_rng(7)
    perm = rng.permutation(len(batch.epj)).astype(np.int32)
    store = np.zeros_like(batch.epj); store[perm] = batch.epj
    shuffled = WalkBatch(store, batch.spj, batch.epi, batch.i_off, perm[batch.id_epj], batch.ej_off, batch.id_spj, batch.sj_off)
    out = {}
    for runs in (1, 0):
        engine.set_option("ep_runs", runs)
        try:
            engine.get_profile(reset=True)
            a = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"], n_walk_limit=16).copy()
            h2d = engine.get_profile()["h2d_bytes"]
            b = engine.calc_force_all_and_write_back(shuffled, prm["eps"], prm["r_out"], prm["G"], n_walk_limit=16).copy()
            c = engine.tree_neighbor_search(batch, n_walk_limit=16).copy()
            out[runs] = (a, b, c, h2d)
        finally:
            engine.set_option("ep_runs", 0)
    for k in range(3):
        assert out[1][k].tobytes() == out[0][k].tobytes()
    assert out[1][3] < 0.8 * out[0][3]                     # fewer bytes cross PCIe
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    assert np.array_equal(out[1][0]["n_ngb"], ref["n_ngb"])
    # recorded dispatches keep their own expanded lists
    L = engine.load()
    engine.set_option("ep_runs", 1)
    engine.check(L.pb_record_begin(), "record")
    engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"], n_walk_limit=16)
    engine.check(L.pb_record_end(), "record")
    engine.calc_force_all_and_write_back(shuffled, prm["eps"], prm["r_out"], prm["G"], n_walk_limit=16)      # overwrites the slots' buffers
    ms = C.c_float(0)
    engine.check(L.pb_replay(2, C.byref(ms), None), "replay")
    engine.set_option("ep_runs", 0)
    assert ms.value > 0


def test_let_send_rows_made_on_the_device():
    """pb_let_gather_epj / pb_let_pack_spj (the LET send rows of a multi-GPU step), with and without raw_upload: the rows
    equal the host packers' bit for bit; with raw_upload a bad index is found by the gather kernel and reported by the
    next call that has synchronised with the device."""
    import torch
    L = engine.load()
    batch, _, prm, _ = hz.kroupa_binary_case(4000)
    n_e, n_s = len(batch.epj), len(batch.spj)
    rng = np.random.default_rng(3)
    idx = rng.integers(0, n_e, size=5000).astype(np.int32)
    sp_rows = batch.spj[rng.integers(0, n_s, size=3000)].copy()
    ref_e = np.zeros((len(idx), 8), dtype=np.float32)
    ref_s = np.zeros((len(sp_rows), 16), dtype=np.float32)
    idx64 = idx.astype(np.int64)
    engine.check(L.pb_pack_epj_host_indexed(batch.epj.ctypes.data, idx64.ctypes.data, len(idx), C.byref(engine.LAYOUT_EPJ), ref_e.ctypes.data), "pack")
    engine.check(L.pb_pack_spj_host(sp_rows.ctypes.data, len(sp_rows), C.byref(engine.LAYOUT_SPJ), ref_s.ctypes.data), "pack")
    d_e = torch.zeros((len(idx), 8), dtype=torch.float32, device="cuda")
    d_s = torch.zeros((len(sp_rows), 16), dtype=torch.float32, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for raw in (0, 1):
        engine.set_option("raw_upload", raw)
        try:
            d_e.zero_(); d_s.zero_(); torch.cuda.synchronize()
            engine.check(L.pb_upload_j(batch.epj.ctypes.data, n_e, C.byref(engine.LAYOUT_EPJ), batch.spj.ctypes.data, n_s, C.byref(engine.LAYOUT_SPJ)), "pb_upload_j")
            engine.check(L.pb_let_gather_epj(idx.ctypes.data, len(idx), d_e.data_ptr()), "pb_let_gather_epj")
            engine.check(L.pb_let_pack_spj(sp_rows.ctypes.data, len(sp_rows), C.byref(engine.LAYOUT_SPJ), d_s.data_ptr()), "pb_let_pack_spj")
            engine.check(L.pb_stream_wait_upload(stream), "pb_stream_wait_upload")
            torch.cuda.synchronize()
            assert d_e.cpu().numpy().tobytes() == ref_e.tobytes(), raw
            assert d_s.cpu().numpy().tobytes() == ref_s.tobytes(), raw
            bad = idx.copy(); bad[17] = n_e + 5
            if raw:
                engine.check(L.pb_let_gather_epj(bad.ctypes.data, len(bad), d_e.data_ptr()), "pb_let_gather_epj")   # accepted: nobody reads it on the host
                engine.check(L.pb_stream_wait_upload(stream), "pb_stream_wait_upload")
                torch.cuda.synchronize()
                with pytest.raises(engine.PbError):
                    engine.check(L.pb_let_gather_epj(idx.ctypes.data, len(idx), d_e.data_ptr()), "pb_let_gather_epj")
                engine.check(L.pb_let_gather_epj(idx.ctypes.data, len(idx), d_e.data_ptr()), "pb_let_gather_epj")   # reported once
            else:
                with pytest.raises(engine.PbError):
                    engine.check(L.pb_let_gather_epj(bad.ctypes.data, len(bad), d_e.data_ptr()), "pb_let_gather_epj")
            torch.cuda.synchronize()
        finally:
            engine.set_option("raw_upload", 0)
