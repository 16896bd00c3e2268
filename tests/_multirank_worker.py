"""Worker of tests/test_multirank.py: world_size-2 (or more) gloo run of the multi-GPU host logic on CPU.

Checks, per rank: (1) the per-step LET all-to-all, carried in the DEVICE j format, lands rows that
are bit-identical to packing the rank's own view of those j (store order == index-list order);
(2) every walk still sees the whole system's mass; (3) forces over local + LET lists (fp64 oracle
as the checker) agree with the single-domain lists to tree-approximation accuracy and neighbour
counts are identical."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from petar_b200 import engine, harness as hz, multigpu  # noqa: E402
from oracle import binding as ob  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 6000
    mass, pos, vel = hz.make_plummer(n)
    prm = hz.petar_auto_params(mass, vel)
    r_in, r_out, rs = hz.particle_rout_rsearch(mass, vel, prm)
    wl = multigpu.build_domain_workload(pos, mass, vel, rs, r_in, r_out, rank, world, dist)
    wl["prm"] = prm
    b = wl["batch"]
    owner = multigpu.domain_split(pos, world)
    assert len(wl["my"]) == (owner == rank).sum() and abs(len(wl["my"]) - n / world) <= 1

    st = multigpu.DomainStepper(wl, rank, world, dist, device=False)
    st.pack_sends()
    st.exchange()
    L = engine.load()
    exp_ep = np.zeros((len(b.epj), 8), dtype=np.float32)
    exp_sp = np.zeros((len(b.spj), 16), dtype=np.float32)
    assert L.pb_pack_epj_host(b.epj.ctypes.data, len(b.epj), C.byref(engine.LAYOUT_EPJ), exp_ep.ctypes.data) == 0
    assert L.pb_pack_spj_host(b.spj.ctypes.data, len(b.spj), C.byref(engine.LAYOUT_SPJ), exp_sp.ctypes.data) == 0
    got_ep, got_sp = st.store_ep.numpy(), st.store_sp.numpy()
    assert wl["n_let_ep"] > 0 and wl["n_let_sp"] > 0
    assert np.array_equal(got_ep[wl["n_loc"]:], exp_ep[wl["n_loc"]:]), "LET EP rows differ from the receiver's packing"
    assert np.array_equal(got_sp[wl["n_nodes"]:], exp_sp[wl["n_nodes"]:]), "LET SP rows differ from the receiver's packing"
    assert st.nccl_bytes_per_step == 32 * sum(st.in_ep) + 64 * sum(st.in_sp)

    mtot = mass.sum()
    for w in range(b.n_walk):
        e = b.id_epj[b.ej_off[w]:b.ej_off[w + 1]]
        s = b.id_spj[b.sj_off[w]:b.sj_off[w + 1]]
        assert abs(b.epj["mass"][e].sum() + b.spj["mass"][s].sum() - mtot) < 1e-12

    # the tree handed to the device-side walk (pb_tree_upload_let): walking it in Python with the element map must
    # give this rank's lists in STORE order (local particles first, LET entries where the all-to-all put them)
    from test_harness import _python_walk
    cells, groups, em = wl["tree_cells"], wl["tree_groups"], wl["elem_map"]
    assert (em < 0).sum() == wl["n_let_sp"] and np.array_equal(np.sort(em[em >= 0]), np.arange(len(b.epj)))
    for g in np.random.default_rng(rank).choice(b.n_walk, min(6, b.n_walk), replace=False):
        ep, sp = _python_walk(cells, groups[g], 0.3, em, len(cells))
        assert np.array_equal(ep, np.sort(b.id_epj[b.ej_off[g]:b.ej_off[g + 1]])), f"EP list of group {g}"
        assert np.array_equal(sp, np.sort(b.id_spj[b.sj_off[g]:b.sj_off[g + 1]])), f"SP list of group {g}"

    f = ob.walks_index(b, prm["eps"], prm["r_out"], prm["G"])
    one, src = hz.build_walk_batch(pos, mass, rs)
    fo = ob.walks_index(one, prm["eps"], prm["r_out"], prm["G"])
    ref = np.zeros(n, dtype=fo.dtype)
    ref[src] = fo
    mine = ref[wl["my"][wl["epi_src"]]]
    assert np.array_equal(f["n_ngb"], mine["n_ngb"])
    err = np.linalg.norm(f["acc"] - mine["acc"], axis=1) / np.linalg.norm(mine["acc"], axis=1)
    assert np.median(err) < 5e-4 and err.max() < 5e-2, (np.median(err), err.max())
    tot = torch.tensor([float(b.n_epi_total)])
    dist.all_reduce(tot)
    assert int(tot.item()) == n
    print(f"rank {rank}/{world} OK n_loc={wl['n_loc']} let_ep={wl['n_let_ep']} let_sp={wl['n_let_sp']} median_err={np.median(err):.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
