"""interaction lists built on the GPU from the uploaded tree (pb_tree_upload /
pb_tree_force) instead of travelling from the host.

* index work is bit-exact: for every i-group the device-built EP and SP lists are the same SETS as
  the host walk's (fp64 geometry without FMA contraction -> identical opening decisions);
* forces through the device-list path meet the same tolerance against the fp64 oracle and the
  neighbour counts are identical;
* at N = 1e6 the whole step (tree upload + device walk + force) is reported next to the host-list path."""
import time

import numpy as np
import pytest

from petar_b200 import engine, harness as hz
from oracle import binding as ob

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("coords0")]   # kernel-level parity: see conftest.coords0


def _case(kind, n):
    batch, _, prm, _ = hz.plummer_case(n) if kind == "plummer" else hz.kroupa_binary_case(n)
    cells, groups = batch.tree.export_tree()
    return batch, prm, cells, groups


@pytest.mark.parametrize("kind,n", [("plummer", 20000), ("kroupa_binaries", 20000)])
def test_device_lists_equal_host_lists_as_sets(kind, n):
    batch, prm, cells, groups = _case(kind, n)
    assert len(groups) == batch.n_walk and np.array_equal(groups["n"], batch.n_epi)
    engine.set_option("tree_batch", 1 << 20)                      # one batch: lists stay readable
    f = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])
    ne, ns, ide, ids = engine.tree_lists(len(groups))
    engine.set_option("tree_batch", 256)
    assert np.array_equal(ne, batch.n_epj) and np.array_equal(ns, batch.n_spj), "list lengths differ from the host walk"
    eo = np.concatenate([[0], np.cumsum(ne)])
    so = np.concatenate([[0], np.cumsum(ns)])
    for g in range(len(groups)):
        assert np.array_equal(np.sort(ide[eo[g]:eo[g + 1]]), batch.id_epj[batch.ej_off[g]:batch.ej_off[g + 1]]), f"EP list of group {g}"
        assert np.array_equal(np.sort(ids[so[g]:so[g + 1]]), np.sort(batch.id_spj[batch.sj_off[g]:batch.sj_off[g + 1]])), f"SP list of group {g}"
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    ea = np.linalg.norm(f["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ep = np.abs((f["pot"] - ref["pot"]) / ref["pot"])
    print(f"[device walk {kind} N={n}] lists identical for {len(groups)} groups; acc rel err median {np.median(ea):.3e} max {ea.max():.3e} | "
          f"pot max {ep.max():.3e} | n_ngb mismatches {(f['n_ngb'] != ref['n_ngb']).sum()}")
    assert np.median(ea) <= 1e-6 and ea.max() <= 1e-4 and np.median(ep) <= 1e-6 and ep.max() <= 1e-4
    assert np.array_equal(f["n_ngb"], ref["n_ngb"])


def test_device_walk_batched_equals_single_batch():
    batch, prm, cells, groups = _case("plummer", 50000)
    engine.set_option("tree_batch", 1 << 20)
    f1 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])
    engine.set_option("tree_batch", 37)                            # many ragged batches cycling over the streams
    f2 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])
    engine.set_option("tree_batch", 256)
    assert np.array_equal(f1["n_ngb"], f2["n_ngb"])
    assert np.abs(f1["acc"] - f2["acc"]).max() <= 2e-6 * np.abs(f1["acc"]).max()
    assert np.abs(f1["pot"] - f2["pot"]).max() <= 2e-6 * np.abs(f1["pot"]).max()
    fh = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"])       # host-list path, same physics
    assert np.array_equal(fh["n_ngb"], f1["n_ngb"])
    assert np.abs(fh["acc"] - f1["acc"]).max() <= 2e-6 * np.abs(f1["acc"]).max()


def test_device_walk_fullsize_report():
    batch, prm, cells, groups = _case("kroupa_binaries", 1000000)
    engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])            # warm-up (allocations)
    engine.get_profile(reset=True)
    t0 = time.perf_counter()
    f = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])
    dt_dev = time.perf_counter() - t0
    prof = engine.get_profile()
    force = np.zeros_like(f)
    tables = engine.make_dispatch_tables(batch, force)
    engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"], force=force, tables=tables)
    t1 = time.perf_counter()
    engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"], force=force, tables=tables)
    dt_host = time.perf_counter() - t1
    inter = sum(batch.interactions())
    print(f"[N=1e6 config 3 stand-in, {len(cells)} cells, {len(groups)} groups] device-walk step {dt_dev * 1e3:.1f} ms "
          f"({inter / dt_dev * 1e-9:.0f} Gint/s, H2D {prof['h2d_bytes'] / 1e6:.0f} MB) vs host-list step {dt_host * 1e3:.1f} ms "
          f"({inter / dt_host * 1e-9:.0f} Gint/s, lists prebuilt; building them on the host took {batch.tree.timing()[1] * 1e3:.0f} ms "
          f"with OpenMP on {__import__('os').cpu_count()} cores); interactions/step {inter:.3e}")
    assert np.array_equal(f["n_ngb"], force["n_ngb"])
    assert np.abs(f["acc"] - force["acc"]).max() <= 2e-6 * np.abs(force["acc"]).max()
    assert prof["n_interaction_ep"] == batch.interactions()[0] and prof["n_interaction_sp"] == batch.interactions()[1]


def _two_domain_case(n=40000):
    """Domain A (x < median) of a Plummer model with the local essential tree it receives from domain B."""
    mass, pos, vel = hz.make_plummer(n)
    prm = hz.petar_auto_params(mass, vel)
    r_in, r_out, rs = hz.particle_rout_rsearch(mass, vel, prm)
    a = pos[:, 0] < np.median(pos[:, 0])
    b = ~a
    ta = hz.TreeHandle(pos[a], mass[a], rs[a])
    tb = hz.TreeHandle(pos[b], mass[b], rs[b])
    ep_idx, sp = tb.make_let(ta.local_boxes())                    # what B owes A: particles near the border, multipoles beyond
    let = dict(pos=pos[b][ep_idx], mass=mass[b][ep_idx], rsearch=rs[b][ep_idx], spj=sp)
    batch, _ = hz.build_walk_batch(pos[a], mass[a], rs[a], vel=vel[a], r_in=r_in[a], r_out=r_out[a], let=let)
    assert len(ep_idx) > 100 and len(sp) > 100 and len(batch.epj) == a.sum() + len(ep_idx)
    return batch, prm


def test_device_walk_with_local_essential_tree():
    """Multi-domain trees: leaves hold EP *and* SP received from other domains; elem_map tells the walk where
    each sorted element is stored.  Lists must still be the host walk's sets; forces meet the tolerance."""
    batch, prm = _two_domain_case()
    cells, groups = batch.tree.export_tree()
    emap = batch.tree.export_elem_map()
    n_let_sp = int((emap < 0).sum())
    assert n_let_sp > 0 and cells["n_let_sp"].sum() == n_let_sp and len(batch.spj) == len(cells) + n_let_sp
    engine.set_option("tree_batch", 1 << 20)
    f = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], elem_map=emap)
    ne, ns, ide, ids = engine.tree_lists(len(groups))
    engine.set_option("tree_batch", 256)
    assert np.array_equal(ne, batch.n_epj) and np.array_equal(ns, batch.n_spj)
    eo, so = np.concatenate([[0], np.cumsum(ne)]), np.concatenate([[0], np.cumsum(ns)])
    saw_let_sp = False
    for g in range(len(groups)):
        assert np.array_equal(np.sort(ide[eo[g]:eo[g + 1]]), np.sort(batch.id_epj[batch.ej_off[g]:batch.ej_off[g + 1]])), f"EP list of group {g}"
        sp_g = np.sort(ids[so[g]:so[g + 1]])
        assert np.array_equal(sp_g, np.sort(batch.id_spj[batch.sj_off[g]:batch.sj_off[g + 1]])), f"SP list of group {g}"
        saw_let_sp |= bool(len(sp_g) and sp_g[-1] >= len(cells))
    assert saw_let_sp
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    ea = np.linalg.norm(f["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    assert np.median(ea) <= 1e-6 and ea.max() <= 1e-4 and np.array_equal(f["n_ngb"], ref["n_ngb"])

    # the j store in another order (as in the multi-GPU step: local particles first, LET entries as they arrive)
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(batch.epj)).astype(np.int32)        # sorted index k lives in store slot perm[k]
    epj_store = np.zeros_like(batch.epj)
    epj_store[perm] = batch.epj
    shuffled = type(batch)(epj_store, batch.spj, batch.epi, batch.i_off, perm[batch.id_epj], batch.ej_off, batch.id_spj, batch.sj_off)
    emap_store = np.where(emap >= 0, perm[np.maximum(emap, 0)], emap).astype(np.int32)
    f2 = engine.tree_force(shuffled, cells, groups, prm["eps"], prm["r_out"], prm["G"], elem_map=emap_store)
    assert np.array_equal(f2["n_ngb"], f["n_ngb"])
    assert np.abs(f2["acc"] - f["acc"]).max() <= 2e-6 * np.abs(f["acc"]).max()


def test_speculative_single_pass_fill_and_its_fallback():
    """From the second tree step on, list space is reserved from the previous step's lengths and the walk fills the
    lists in ONE pass; a list that outgrows its reservation must be caught and redone exactly."""
    batch, prm, cells, groups = _case("kroupa_binaries", 20000)
    engine.set_option("tree_spec", 0)
    ref = {th: engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=th).copy() for th in (0.5, 0.3)}
    engine.set_option("tree_spec", 1)
    f1 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.5).copy()     # no history yet or stale: exact path
    f2 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.5).copy()     # speculative, fits
    ne2, ns2, ide2, ids2 = engine.tree_lists(len(groups))
    assert f1.tobytes() == ref[0.5].tobytes() and f2.tobytes() == ref[0.5].tobytes()
    f3 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.3).copy()     # lists ~2x longer: reservation overflows
    ne3, ns3, ide3, ids3 = engine.tree_lists(len(groups))
    assert ns3.sum() > 1.3 * ns2.sum()
    assert f3.tobytes() == ref[0.3].tobytes()
    f4 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.3).copy()     # speculative again
    ne4, ns4, ide4, ids4 = engine.tree_lists(len(groups))
    assert f4.tobytes() == ref[0.3].tobytes() and np.array_equal(ne4, ne3) and np.array_equal(ide4, ide3) and np.array_equal(ids4, ids3)


def test_tree_from_pinned_staging_buffers():
    batch, prm, cells, groups = _case("plummer", 20000)
    f0 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"]).copy()
    sc, sg = batch.tree.export_tree(out=engine.tree_stage(len(cells), len(groups)))
    assert sc.tobytes() == cells.tobytes() and sg.tobytes() == groups.tobytes()
    f1 = engine.tree_force(batch, sc, sg, prm["eps"], prm["r_out"], prm["G"]).copy()
    assert f1.tobytes() == f0.tobytes()


# ---- device-resident step: i-particles from the j store, plan and forces on the GPU (pb_tree_force_resident) ----------
@pytest.mark.parametrize("kind,n", [("plummer", 20000), ("kroupa_binaries", 20000)])
def test_resident_step_equals_host_planned_step(kind, n):
    """Same lists, same kernels, the plan made on the device: forces agree with the host-planned tree step to summation-
    order level (the walk origins differ in the last bits: mean by another summation order), counts exactly; the
    oracle tolerance holds; a second (speculative, no host round trip) step reproduces the first bit for bit."""
    batch, prm, cells, groups = _case(kind, n)
    fh = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"]).copy()
    f1 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], resident=True).copy()
    f2 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], resident=True).copy()
    tl = engine.tree_timeline()
    print(f"[resident {kind} N={n}] device timeline ms {tl}")
    assert np.array_equal(f1["n_ngb"], fh["n_ngb"])
    assert np.abs(f1["acc"] - fh["acc"]).max() <= 2e-6 * np.abs(fh["acc"]).max()
    assert np.abs(f1["pot"] - fh["pot"]).max() <= 2e-6 * np.abs(fh["pot"]).max()
    assert f2.tobytes() == f1.tobytes()
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    ea = np.linalg.norm(f1["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
    ep = np.abs((f1["pot"] - ref["pot"]) / ref["pot"])
    assert np.median(ea) <= 1e-6 and ea.max() <= 1e-4 and np.median(ep) <= 1e-6 and ep.max() <= 1e-4
    assert np.array_equal(f1["n_ngb"], ref["n_ngb"])


def test_resident_step_reservation_overflow_falls_back():
    """theta 0.5 -> 0.3 roughly doubles the SP lists: the second step's reservations (lists, tasks, partial sums) are
    too small, the device detects it without running anything past them, and the step is repeated exactly."""
    batch, prm, cells, groups = _case("kroupa_binaries", 20000)
    ref03 = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.3).copy()
    a = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.5, resident=True).copy()
    b = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.5, resident=True).copy()
    c = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.3, resident=True).copy()   # overflow -> exact repeat
    d = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], theta=0.3, resident=True).copy()   # speculative again
    assert a.tobytes() == b.tobytes() and c.tobytes() == d.tobytes()
    assert np.array_equal(c["n_ngb"], ref03["n_ngb"])
    assert np.abs(c["acc"] - ref03["acc"]).max() <= 2e-6 * np.abs(ref03["acc"]).max()
    assert not np.array_equal(a["acc"], c["acc"])


def test_resident_step_with_local_essential_tree():
    batch, prm = _two_domain_case()
    cells, groups = batch.tree.export_tree()
    emap = batch.tree.export_elem_map()
    # store order of the multi-GPU step: local particles first (in i-group order), LET entries behind them
    n_loc = batch.n_epi_total
    epj_src_is_local = np.zeros(len(batch.epj), bool)
    fh = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], elem_map=emap).copy()
    # sorted-order store: group g's particles are NOT a contiguous slice of it in a tree with LET elements, so build the
    # store the resident step expects: slot k < n_loc = k-th i-particle, LET EP behind
    ids_i = batch.epi["id"]
    pos_of_id = {int(i): k for k, i in enumerate(batch.epj["id"])}
    loc_sorted = np.array([pos_of_id[int(i)] for i in ids_i], dtype=np.int64)          # sorted index of every i-particle
    is_loc = np.zeros(len(batch.epj), bool); is_loc[loc_sorted] = True
    let_sorted = np.nonzero(~is_loc)[0]
    perm = np.empty(len(batch.epj), dtype=np.int32)                                    # sorted index k lives in store slot perm[k]
    perm[loc_sorted] = np.arange(n_loc, dtype=np.int32)
    perm[let_sorted] = n_loc + np.arange(len(let_sorted), dtype=np.int32)
    epj_store = np.zeros_like(batch.epj); epj_store[perm] = batch.epj
    store = type(batch)(epj_store, batch.spj, batch.epi, batch.i_off, perm[batch.id_epj], batch.ej_off, batch.id_spj, batch.sj_off)
    emap_store = np.where(emap >= 0, perm[np.maximum(emap, 0)], emap).astype(np.int32)
    g2 = groups.copy(); g2["first"] = batch.i_off[:-1]                                 # = store slot of each group's first particle
    fr = engine.tree_force(store, cells, g2, prm["eps"], prm["r_out"], prm["G"], elem_map=emap_store, resident=True)
    assert np.array_equal(fr["n_ngb"], fh["n_ngb"])
    assert np.abs(fr["acc"] - fh["acc"]).max() <= 2e-6 * np.abs(fh["acc"]).max()


def test_compact_walk_records_give_identical_lists_in_identical_order():
    """walk_compact = 1 (default) classifies on 64-byte fp32 records rounded outward and re-checks undecided cells on the
    fp64 record: the lists must be the legacy all-fp64 walk's lists entry for entry, for a plain tree, for binaries with
    their tight clumps, far from the origin (large |coordinate| = large fp32 rounding margin) and with a LET tree."""
    def lists(batch, cells, groups, prm, emap=None):
        engine.set_option("tree_batch", 1 << 20)
        try:
            f = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], elem_map=emap).copy()
            return f, engine.tree_lists(len(groups))
        finally:
            engine.set_option("tree_batch", 256)

    cases = []
    for kind in ("plummer", "kroupa_binaries"):
        batch, prm, cells, groups = _case(kind, 20000)
        cases.append((kind, batch, cells, groups, prm, None))
    # the same Plummer model pushed far from the origin: |x| ~ 1e3 makes the fp32 margin comparable to small cells
    mass, pos, vel = hz.make_plummer(20000)
    prm = hz.petar_auto_params(mass, vel)
    r_in, r_out, rs = hz.particle_rout_rsearch(mass, vel, prm)
    far, _ = hz.build_walk_batch(pos + np.array([1.0e3, -2.0e3, 5.0e2]), mass, rs, vel=vel, r_in=r_in, r_out=r_out)
    c2, g2 = far.tree.export_tree()
    cases.append(("plummer shifted by 1e3", far, c2, g2, prm, None))
    lb, lprm = _two_domain_case()
    lc, lg = lb.tree.export_tree()
    cases.append(("two domains (LET)", lb, lc, lg, lprm, lb.tree.export_elem_map()))
    for name, batch, cells, groups, prm, emap in cases:
        engine.set_option("walk_compact", 0)
        try:
            f0, l0 = lists(batch, cells, groups, prm, emap)
        finally:
            engine.set_option("walk_compact", 1)
        f1, l1 = lists(batch, cells, groups, prm, emap)
        for a, b in zip(l0, l1):
            assert np.array_equal(a, b), name
        assert f0.tobytes() == f1.tobytes(), name
        print(f"[compact walk, {name}] {len(groups)} groups, {int(l1[0].sum())} EP + {int(l1[1].sum())} SP list entries identical")


def test_fused_reduction_and_direct_result_write_are_bitwise_equivalent():
    """Options fuse_reduce (the force kernel adds up finished i-blocks and writes the forces into page-locked host memory) and
    raw_result (... into the caller's own array): the same partial sums are added in the same chunk order whoever delivers
    the last one, so every combination returns the same bytes as reduction kernel + D2H copy; both kernel variants; ragged
    walks; repeated steps (the chunk counters return to zero)."""
    batch, prm, cells, groups = _case("kroupa_binaries", 20000)
    out = {}
    f = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)
    try:
        for ws, sp2i in ((1, 1), (1, 0)):
            for fuse, raw in ((0, 0), (1, 0), (1, 1)):
                engine.set_option("ws", ws); engine.set_option("sp2i", sp2i)
                engine.set_option("fuse_reduce", fuse); engine.set_option("raw_result", raw)
                for _ in range(3):                                  # exact first step, then speculative ones
                    f[:] = 0
                    engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], force=f, resident=True)
                out[(ws, sp2i, fuse, raw)] = f.copy()
    finally:
        engine.set_option("ws", 1); engine.set_option("sp2i", 1); engine.set_option("fuse_reduce", 1); engine.set_option("raw_result", 0)
    for ws, sp2i in ((1, 1), (1, 0)):
        base = out[(ws, sp2i, 0, 0)]
        assert np.abs(base["acc"]).max() > 0
        assert out[(ws, sp2i, 1, 0)].tobytes() == base.tobytes()
        assert out[(ws, sp2i, 1, 1)].tobytes() == base.tobytes()
    ref = ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"])
    assert np.array_equal(out[(1, 1, 1, 1)]["n_ngb"], ref["n_ngb"])
