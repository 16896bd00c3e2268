"""CPU tests of the walk-list harness: the lists it emits have the properties FDPS's
QuadrupoleWithSymmetrySearch walk guarantees (see petar_b200/harness/tree_walk.cpp header)."""
import numpy as np
import pytest

from petar_b200 import harness as hz
from oracle import binding as ob


@pytest.fixture(scope="module")
def case():
    batch, epi_src, prm, (mass, pos, vel) = hz.plummer_case(3000)
    return batch, epi_src, prm, mass, pos, vel


def test_plummer_recipe_matches_oracle_bitwise():
    m, p, v = hz.make_plummer(1500)
    m2, p2, v2 = ob.make_plummer(1500)
    assert np.array_equal(m, m2) and np.array_equal(p, p2) and np.array_equal(v, v2)
    # Henon units: total energy -1/4 => virial: 2T ~ 0.5, |W| ~ 0.5
    ke = 0.5 * (m[:, None] * v * v).sum()
    assert 0.2 < ke < 0.3


def test_petar_auto_params_match_survey():
    m, p, v = hz.make_plummer(100000)
    prm = hz.petar_auto_params(m, v)
    assert abs(prm["vel_disp"] - 0.41) < 0.01          # sigma_1D of a virialised Plummer sphere
    assert abs(prm["r_out"] - 4.3e-3) < 2e-4
    assert prm["dt_soft"] == 2.0 ** -10
    assert hz.regular_time_step(0.3) == 0.25 and hz.regular_time_step(3.0) == 2.0 and hz.regular_time_step(1.0) == 1.0


def test_walks_partition_the_particles(case):
    batch, epi_src, *_ = case
    assert sorted(epi_src.tolist()) == list(range(3000))
    assert batch.n_epi.max() <= hz.N_GROUP_LIMIT
    assert batch.n_epi_total == 3000


def test_every_walk_sees_all_the_mass(case):
    """EP list + SP list of a walk tile the whole system exactly once: masses add up to M."""
    batch, *_ = case
    mtot = batch.epj["mass"].sum()
    for w in range(batch.n_walk):
        e = batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]]
        s = batch.id_spj[batch.sj_off[w]:batch.sj_off[w + 1]]
        assert len(np.unique(e)) == len(e)
        mw = batch.epj["mass"][e].sum() + batch.spj["mass"][s].sum()
        assert abs(mw - mtot) < 1e-12
        assert np.all(np.diff(e) > 0)          # Morton order, like FDPS: runs of consecutive indices


def test_symmetric_search_no_neighbour_inside_a_superparticle(case):
    """Any j within max(rs_i, rs_j) of an i of the group is in that group's EP list, so the neighbour
    count over the walk list equals the count over ALL particles."""
    batch, *_ = case
    f_walk = ob.walks_index(batch, 0.0, 1e-3, 1.0)
    nb_all = ob.search_neighbor(batch.epi, batch.epj)
    assert np.array_equal(f_walk["n_ngb"], nb_all["n_ngb"])
    assert f_walk["n_ngb"].min() >= 1            # every particle counts itself


def test_tree_force_close_to_direct_sum(case):
    """theta = 0.3 quadrupole: walk-list force within ~1e-3 of the direct sum (with the same cutoff)."""
    batch, _, prm, *_ = case
    f_walk = ob.walks_index(batch, prm["eps"], prm["r_out"], 1.0)
    f_dir = ob.force_epep(batch.epi, batch.epj, prm["eps"], prm["r_out"], 1.0)
    err = np.linalg.norm(f_walk["acc"] - f_dir["acc"], axis=1) / np.linalg.norm(f_dir["acc"], axis=1)
    assert np.median(err) < 2e-4 and err.max() < 2e-2
    assert np.abs((f_walk["pot"] - f_dir["pot"]) / f_dir["pot"]).max() < 1e-4


def test_superparticle_moments_are_about_the_centre_of_mass(case):
    batch, *_ = case
    root = batch.spj[0]                                   # node 0 = whole system
    m, x = batch.epj["mass"], batch.epj["pos"]
    cm = (m[:, None] * x).sum(0) / m.sum()
    assert np.allclose(root["pos"], cm, atol=1e-12) and np.isclose(root["mass"], m.sum())
    d = x - cm
    q = np.array([(m * d[:, 0] * d[:, 0]).sum(), (m * d[:, 1] * d[:, 1]).sum(), (m * d[:, 2] * d[:, 2]).sum(),
                  (m * d[:, 0] * d[:, 1]).sum(), (m * d[:, 0] * d[:, 2]).sum(), (m * d[:, 1] * d[:, 2]).sum()])
    assert np.allclose(root["quad"], q, rtol=1e-9, atol=1e-12)


def test_let_two_domains_reproduce_single_domain_lists():
    """Split the system in two boxes, exchange LET (EP near, SP far), rebuild: each domain's walks must
    again see all the mass, keep every neighbour as EP, and give forces close to the one-domain ones."""
    mass, pos, vel = hz.make_plummer(4000)
    prm = hz.petar_auto_params(mass, vel)
    _, _, rs = hz.particle_rout_rsearch(mass, vel, prm)
    left = pos[:, 0] < np.median(pos[:, 0])
    doms = [np.nonzero(left)[0], np.nonzero(~left)[0]]
    trees = [hz.TreeHandle(pos[d], mass[d], rs[d]) for d in doms]
    boxes = [t.local_boxes() for t in trees]
    f_ref_batch, src_ref = hz.build_walk_batch(pos, mass, rs)
    f_ref = ob.walks_index(f_ref_batch, 0.0, prm["r_out"], 1.0)
    ref_by_particle = np.zeros(4000, dtype=f_ref.dtype)
    ref_by_particle[src_ref] = f_ref
    for me, other in ((0, 1), (1, 0)):
        ep_idx, sp = trees[other].make_let(boxes[me])
        assert 0 < len(ep_idx) < len(doms[other]) and len(sp) > 0
        sent = doms[other][ep_idx]
        assert abs(mass[sent].sum() + sp["mass"].sum() - mass[doms[other]].sum()) < 1e-12
        let = dict(pos=pos[sent], mass=mass[sent], rsearch=rs[sent], spj=sp)
        batch, src = hz.build_walk_batch(pos[doms[me]], mass[doms[me]], rs[doms[me]], let=let)
        for w in range(batch.n_walk):
            e = batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]]
            s = batch.id_spj[batch.sj_off[w]:batch.sj_off[w + 1]]
            assert abs(batch.epj["mass"][e].sum() + batch.spj["mass"][s].sum() - mass.sum()) < 1e-12
        f = ob.walks_index(batch, 0.0, prm["r_out"], 1.0)
        ref = ref_by_particle[doms[me][src]]
        assert np.array_equal(f["n_ngb"], ref["n_ngb"])
        err = np.linalg.norm(f["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
        assert np.median(err) < 5e-4 and err.max() < 5e-2


# ---- the exported tree (pb_tree_cell / pb_tree_group / elem_map) reproduces the host walk's lists ----------
def _python_walk(cells, grp, theta, elem_map, n_cells):
    """The opening rule of include/petar_b200.h ("device-side interaction lists"), breadth-first, in fp64."""
    inv2 = 1.0 / (theta * theta)
    ep, sp, frontier = [], [], [0]
    while frontier:
        nxt = []
        for c in frontier:
            cell = cells[c]
            if cell["n"] == 0:
                continue
            d = np.maximum(np.maximum(grp["in_lo"] - cell["cm"], cell["cm"] - grp["in_hi"]), 0.0)
            far = (d * d).sum() > cell["len"] * cell["len"] * inv2
            touch = (np.all(grp["out_lo"] <= cell["in_hi"]) and np.all(grp["out_hi"] >= cell["in_lo"])) or \
                    (np.all(cell["out_lo"] <= grp["in_hi"]) and np.all(cell["out_hi"] >= grp["in_lo"]))
            if far and not touch:
                sp.append(c)
            elif cell["leaf"]:
                for k in range(cell["first"], cell["first"] + cell["n"]):
                    m = k if elem_map is None else int(elem_map[k])
                    if m >= 0:
                        ep.append(m)
                    else:
                        sp.append(n_cells + ~m)
            else:
                nxt.extend(int(x) for x in cell["child"] if x >= 0)
        frontier = nxt
    return np.sort(ep), np.sort(sp)


def _check_exported_tree(batch, elem_map):
    cells, groups = batch.tree.export_tree()
    assert len(groups) == batch.n_walk
    rng = np.random.default_rng(0)
    for g in rng.choice(batch.n_walk, min(12, batch.n_walk), replace=False):
        ep, sp = _python_walk(cells, groups[g], 0.3, elem_map, len(cells))
        assert np.array_equal(ep, np.sort(batch.id_epj[batch.ej_off[g]:batch.ej_off[g + 1]])), f"EP list of group {g}"
        assert np.array_equal(sp, np.sort(batch.id_spj[batch.sj_off[g]:batch.sj_off[g + 1]])), f"SP list of group {g}"
    return cells


def test_exported_tree_single_domain():
    batch, _, prm, _ = hz.plummer_case(6000)
    cells = _check_exported_tree(batch, None)
    assert cells["n_let_sp"].sum() == 0
    assert np.array_equal(batch.tree.export_elem_map(), np.arange(len(batch.epj)))       # identity without LET


def test_exported_tree_with_local_essential_tree():
    mass, pos, vel = hz.make_plummer(8000)
    prm = hz.petar_auto_params(mass, vel)
    _, _, rs = hz.particle_rout_rsearch(mass, vel, prm)
    a = pos[:, 0] < np.median(pos[:, 0])
    ta, tb = hz.TreeHandle(pos[a], mass[a], rs[a]), hz.TreeHandle(pos[~a], mass[~a], rs[~a])
    ep_idx, sp = tb.make_let(ta.local_boxes())
    let = dict(pos=pos[~a][ep_idx], mass=mass[~a][ep_idx], rsearch=rs[~a][ep_idx], spj=sp)
    batch, _ = hz.build_walk_batch(pos[a], mass[a], rs[a], let=let)
    em = batch.tree.export_elem_map()
    cells = _check_exported_tree(batch, em)
    assert (em < 0).sum() == len(sp) == cells["n_let_sp"].sum() and len(batch.spj) == len(cells) + len(sp)
    assert np.array_equal(np.sort(em[em >= 0]), np.arange(len(batch.epj)))                # every EP exactly once
    assert np.array_equal(np.sort(~em[em < 0]), np.arange(len(sp)))                       # every LET SP exactly once
