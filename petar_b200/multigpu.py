"""Multi-GPU form of the hot path: one rank per GPU, spatial domain decomposition, and the
local-essential-tree (LET) exchange carried by one NCCL all-to-all per tree step.

In PeTar the decomposition and the LET live in FDPS (``dinfo.decomposeDomainAll``, reference
``src/petar.hpp:1739-1741``; LET inside ``calcForceAllAndWriteBack*``, profiled as ``Make_LET`` /
``Ex_LET``, ``src/profile.hpp:286-289``) and travel over MPI as fp64 ``EPJSoft`` (120 B) /
``SPJQuadrupoleInAndOut`` (80 B).  Here every rank

1. packs the LET entries it owes each peer in the library's DEVICE j format (32 B EP / 64 B SP),
2. sends them with ``all_to_all_single`` (NCCL over NVLink) straight into the receivers' device
   j stores, behind the locally uploaded part — no host round trip, no repack on the receiver,
3. publishes the store to the dispatch streams and runs its walks.  Forces need no reduction.

Which entries go where (the LET *selection*) and the walk lists over local + LET j are FDPS's job;
the harness (``harness/tree_walk.cpp``) stands in for it, once, outside the timed region.
"""
import ctypes as C
import os

import numpy as np

from . import engine, harness as hz
from .types import EPISoft, EPJSoft, SPJQuad, ForceSoft
from .walks import WalkBatch


def domain_split(pos, world):
    """Recursive coordinate bisection into `world` boxes of equal particle count (x, y, z cycling):
    the multisection layout FDPS uses for 2/4/8 ranks.  Returns the owner rank of every particle."""
    owner = np.zeros(len(pos), dtype=np.int32)

    def rec(idx, r0, nparts, axis):
        if nparts == 1:
            owner[idx] = r0
            return
        nl = nparts // 2
        order = idx[np.argsort(pos[idx, axis], kind="stable")]
        cut = len(order) * nl // nparts
        rec(order[:cut], r0, nl, (axis + 1) % 3)
        rec(order[cut:], r0 + nl, nparts - nl, (axis + 1) % 3)

    rec(np.arange(len(pos)), 0, world, 0)
    return owner


def _torch_dev(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def exchange_rows(dist, rows_per_dest, width, dtype):
    """all-to-all-v of row blocks: rows_per_dest[q] is an (n_q, width) array for rank q.
    Returns (list of received (m_q, width) arrays by source rank, recv counts)."""
    import torch
    world = dist.get_world_size()
    dev = _torch_dev(dist)
    cnt_in = torch.tensor([len(r) for r in rows_per_dest], dtype=torch.int64, device=dev)
    cnt_out = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(cnt_out, cnt_in)
    cnt_out_l = [int(x) for x in cnt_out.tolist()]
    send = np.concatenate([np.asarray(r, dtype=dtype).reshape(-1, width) for r in rows_per_dest]) if world else np.zeros((0, width), dtype)
    tsend = torch.from_numpy(np.ascontiguousarray(send)).to(dev)
    trecv = torch.empty((sum(cnt_out_l), width), dtype=tsend.dtype, device=dev)
    dist.all_to_all_single(trecv, tsend, cnt_out_l, [len(r) for r in rows_per_dest])
    out = trecv.cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(cnt_out_l)])
    return [out[offs[q]:offs[q + 1]] for q in range(world)], cnt_out_l


def build_domain_workload(pos, mass, vel, rs, r_in, r_out, rank, world, dist=None, theta=hz.THETA,
                          n_leaf_limit=hz.N_LEAF_LIMIT, n_group_limit=hz.N_GROUP_LIMIT, ptype=None):
    """This rank's domain, its LET send plan, and its walk batch over local + received LET j.

    j-store order on every rank: EP = [local particles in local Morton order | LET EP by source
    rank], SP = [cells of the global tree | LET SP by source rank] — exactly where the per-step
    all-to-all lands them, so the index lists refer to store slots directly."""
    if dist is None:
        import torch.distributed as dist
    owner = domain_split(pos, world)
    my = np.nonzero(owner == rank)[0]
    # local Morton order first, so the local part of the store keeps tree locality
    t0 = hz.TreeHandle(pos[my], mass[my], rs[my], None, theta, n_leaf_limit, n_group_limit)
    my = my[t0.export()[0]]
    t0.close()
    lpos, lmass, lrs, lvel = pos[my], mass[my], rs[my], vel[my]
    tloc = hz.TreeHandle(lpos, lmass, lrs, None, theta, n_leaf_limit, n_group_limit)

    import torch
    dev = _torch_dev(dist)
    boxes = torch.empty(world * 12, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(boxes, torch.from_numpy(tloc.local_boxes()).to(dev))
    boxes = boxes.cpu().numpy().reshape(world, 12)

    send_ep_idx, send_sp = [], []
    for q in range(world):
        if q == rank:
            send_ep_idx.append(np.zeros(0, dtype=np.int32)); send_sp.append(np.zeros(0, dtype=SPJQuad))
        else:
            e, s = tloc.make_let(boxes[q])
            send_ep_idx.append(e); send_sp.append(s)

    # fp64 LET exchange for the (untimed) global-tree build: what FDPS's MPI exchange carries
    # (sixth column: the particle's global index, so that checkers can look up its other attributes)
    ep_rows = [np.column_stack([lpos[e], lmass[e], lrs[e], my[e].astype(np.float64)]) if len(e) else np.zeros((0, 6)) for e in send_ep_idx]
    sp_rows = [s.view(np.float64).reshape(-1, 10) for s in send_sp]
    recv_ep, recv_ep_cnt = exchange_rows(dist, ep_rows, 6, np.float64)
    recv_sp, recv_sp_cnt = exchange_rows(dist, sp_rows, 10, np.float64)
    let_ep = np.concatenate(recv_ep) if recv_ep else np.zeros((0, 6))
    let_sp = np.ascontiguousarray(np.concatenate(recv_sp)).view(SPJQuad).reshape(-1)
    let = dict(pos=let_ep[:, 0:3], mass=let_ep[:, 3], rsearch=let_ep[:, 4], spj=let_sp)

    t = hz.TreeHandle(lpos, lmass, lrs, let, theta, n_leaf_limit, n_group_limit)
    epj_src, epi_src, spj_sorted, i_off, ej_off, sj_off, id_epj, id_spj = t.export()
    n_loc, n_let_ep, n_nodes, n_let_sp = len(my), len(let_ep), t.n_nodes, t.n_let_sp
    assert n_let_sp == len(let_sp) and t.n_epj == n_loc + n_let_ep

    epj = np.zeros(n_loc + n_let_ep, dtype=EPJSoft)           # store (= source) order
    epj["id"][:n_loc] = my + 1
    epj["mass"] = np.concatenate([lmass, let["mass"]])
    epj["pos"] = np.concatenate([lpos, let["pos"]])
    epj["r_search"] = np.concatenate([lrs, let["rsearch"]])
    epj["vel"][:n_loc] = lvel
    epj["r_in"][:n_loc] = r_in[my]; epj["r_out"][:n_loc] = r_out[my]
    epj["r_scale_next"] = 1.0
    epj["rank_org"] = rank
    id_epj_store = epj_src[id_epj].astype(np.int32)

    spj = np.concatenate([spj_sorted[:n_nodes], let_sp])
    sp_src = t.let_sp_src()
    id_spj_store = np.where(id_spj < n_nodes, id_spj, n_nodes + sp_src[np.maximum(id_spj - n_nodes, 0)]).astype(np.int32) if n_let_sp else id_spj

    epi = np.zeros(t.n_epi, dtype=EPISoft)
    epi["id"] = my[epi_src] + 1
    epi["pos"] = lpos[epi_src]
    epi["r_search"] = lrs[epi_src]
    epi["rank_org"] = rank
    epi["type"] = 1 if ptype is None else np.asarray(ptype)[my[epi_src]]
    batch = WalkBatch(epj, spj, epi, i_off, id_epj_store, ej_off, id_spj_store, sj_off)
    batch.tree = t
    # for the device-side walk (pb_tree_upload_let): the global tree and where each of its sorted elements is stored
    cells, groups = t.export_tree()
    em = t.export_elem_map()
    elem_map = np.where(em >= 0, epj_src[np.maximum(em, 0)], ~sp_src[np.maximum(~em, 0)] if n_let_sp else em).astype(np.int32)
    return dict(batch=batch, my=my, epi_src=epi_src, n_loc=n_loc, n_nodes=n_nodes, n_let_ep=n_let_ep, n_let_sp=n_let_sp,
                send_ep_idx=send_ep_idx, send_sp=send_sp, recv_ep_cnt=recv_ep_cnt, recv_sp_cnt=recv_sp_cnt, let=let, local_tree=tloc,
                tree_cells=cells, tree_groups=groups, elem_map=elem_map,
                store_gid=np.concatenate([my, let_ep[:, 5].astype(np.int64)]))


class DomainStepper:
    """One tree step of the hot path on one rank of a multi-GPU run (see module docstring)."""

    def __init__(self, wl, rank, world, dist, device=True):
        import torch
        self.torch, self.dist, self.wl, self.rank, self.world, self.device = torch, dist, wl, rank, world, device
        self.batch, self.prm = wl["batch"], wl["prm"]
        b = self.batch
        self.n_loc, self.n_nodes = wl["n_loc"], wl["n_nodes"]
        self.send_idx = np.concatenate(wl["send_ep_idx"]).astype(np.int64)
        self.send_sp = np.concatenate(wl["send_sp"])
        self.in_ep = [len(x) for x in wl["send_ep_idx"]]
        self.in_sp = [len(x) for x in wl["send_sp"]]
        self.out_ep, self.out_sp = wl["recv_ep_cnt"], wl["recv_sp_cnt"]
        self.nccl_bytes_per_step = 32 * len(self.send_idx) + 64 * len(self.send_sp)
        self.host_s, self.n_steps = {}, 0          # wall-clock of the step's host phases (seconds, accumulated)
        self.host_dw, self.n_dw = {}, 0
        self.trace = {} if os.environ.get("PETAR_B200_TRACE_HOST") else None   # finer host timers of the LET exchange
        self.send_idx32 = self.send_idx.astype(np.int32)            # LET EP rows as store slots (local particles come first)
        self.L = engine.load()
        if device:
            # several ranks share the node's host cores: this rank's own j-particles travel raw and are packed on the GPU
            engine.set_option("raw_upload", int(os.environ.get("PETAR_B200_RAW_UPLOAD", "1")))
            engine.set_option("raw_result", int(os.environ.get("PETAR_B200_RAW_RESULT", "1")))
        pin = device
        self.h_send_ep = torch.empty((len(self.send_idx), 8), dtype=torch.float32, pin_memory=pin)
        self.h_send_sp = torch.empty((len(self.send_sp), 16), dtype=torch.float32, pin_memory=pin)
        if device:
            self.d_send_ep = torch.empty_like(self.h_send_ep, device="cuda")
            self.d_send_sp = torch.empty_like(self.h_send_sp, device="cuda")
            self._bind_store()
        else:
            self.store_ep = torch.zeros((len(b.epj), 8), dtype=torch.float32)
            self.store_sp = torch.zeros((len(b.spj), 16), dtype=torch.float32)

    # the library owns the j store; torch only sees it through __cuda_array_interface__
    def _bind_store(self):
        b, torch = self.batch, self.torch
        pe, ps = C.c_void_p(0), C.c_void_p(0)
        engine.check(self.L.pb_reserve_j(len(b.epj), len(b.spj), C.byref(pe), C.byref(ps)), "pb_reserve_j")
        self._ptrs = (pe.value, ps.value)

        class _Raw:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

        self.store_ep = torch.as_tensor(_Raw(pe.value, (len(b.epj), 8)), device="cuda")
        self.store_sp = torch.as_tensor(_Raw(ps.value, (max(len(b.spj), 1), 16)), device="cuda")[:len(b.spj)]

    def pack_sends(self):
        """Host packing of the LET send rows (EP gathered from the local particles, SP from the local tree's multipoles) —
        the CPU (gloo) path and the reference for the device gather; the GPU steps pack only the SP rows here."""
        b = self.batch
        if len(self.send_idx):                                          # gather from the local particles while packing
            engine.check(self.L.pb_pack_epj_host_indexed(b.epj.ctypes.data, self.send_idx.ctypes.data, len(self.send_idx),
                                                         C.byref(engine.LAYOUT_EPJ), self.h_send_ep.data_ptr()), "pb_pack_epj_host_indexed")
        self.pack_send_sp()

    def pack_send_sp(self):
        if len(self.send_sp):
            engine.check(self.L.pb_pack_spj_host(self.send_sp.ctypes.data, len(self.send_sp), C.byref(engine.LAYOUT_SPJ),
                                                 self.h_send_sp.data_ptr()), "pb_pack_spj_host")

    def exchange(self):
        """LET all-to-all in the device j format; lands behind the local part of the store.  CPU (gloo) form, and the
        round-1 GPU form with host-packed EP rows (kept for A/B: option host_let_pack)."""
        torch, dist = self.torch, self.dist
        if self.device:
            self.d_send_ep.copy_(self.h_send_ep, non_blocking=True)
            self.d_send_sp.copy_(self.h_send_sp, non_blocking=True)
            se, ss = self.d_send_ep, self.d_send_sp
        else:
            se, ss = self.h_send_ep, self.h_send_sp
        dist.all_to_all_single(self.store_ep[self.n_loc:], se, self.out_ep, self.in_ep)
        dist.all_to_all_single(self.store_sp[self.n_nodes:], ss, self.out_sp, self.in_sp)

    def exchange_device(self):
        """The LET exchange of the GPU steps.  EP rows are gathered ON THE DEVICE from the local part of the j store
        (the host ships 4-byte store slots, not packed 32-byte rows); SP rows are the local tree's multipoles, which
        exist on the host only: pb_let_pack_spj copies them as they are and packs them on the device (raw_upload) or packs
        them into pinned staging.  One NCCL all-to-all per kind writes straight into the peers' j stores behind their
        locally uploaded part."""
        import time
        torch, dist, L = self.torch, self.dist, self.L
        tr = self.trace
        t = [time.perf_counter()] if tr is not None else None
        if len(self.send_idx32):
            engine.check(L.pb_let_gather_epj(self.send_idx32.ctypes.data, len(self.send_idx32), self.d_send_ep.data_ptr()), "pb_let_gather_epj")
        if t: t.append(time.perf_counter())
        if len(self.send_sp):
            engine.check(L.pb_let_pack_spj(self.send_sp.ctypes.data, len(self.send_sp), C.byref(engine.LAYOUT_SPJ), self.d_send_sp.data_ptr()), "pb_let_pack_spj")
        if t: t.append(time.perf_counter())
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        engine.check(L.pb_stream_wait_upload(stream), "pb_stream_wait_upload")     # local j copied, EP rows gathered
        if t: t.append(time.perf_counter())
        dist.all_to_all_single(self.store_ep[self.n_loc:], self.d_send_ep, self.out_ep, self.in_ep)
        if t: t.append(time.perf_counter())
        dist.all_to_all_single(self.store_sp[self.n_nodes:], self.d_send_sp, self.out_sp, self.in_sp)
        if t: t.append(time.perf_counter())
        engine.check(L.pb_publish_j(stream), "pb_publish_j")
        if t:
            t.append(time.perf_counter())
            for k, name in enumerate(("gather_ep", "pack_sp", "wait_upload", "all_to_all_ep", "all_to_all_sp", "publish")):
                tr[name] = tr.get(name, 0.0) + t[k + 1] - t[k]
            tr["n"] = tr.get("n", 0) + 1

    def _stage_tree(self):
        wl = self.wl
        if not getattr(self, "_tree_staged", False):          # the tree lives in the library's pinned staging buffers
            sc, sg = engine.tree_stage(len(wl["tree_cells"]), len(wl["tree_groups"]))
            sc[:] = wl["tree_cells"]; sg[:] = wl["tree_groups"]
            wl["tree_cells"], wl["tree_groups"], self._tree_staged = sc, sg, True
        return wl["tree_cells"], wl["tree_groups"], wl["elem_map"]

    def upload_local_j(self):
        b, L = self.batch, self.L
        pe, ps = C.c_void_p(0), C.c_void_p(0)
        engine.check(L.pb_reserve_j(len(b.epj), len(b.spj), C.byref(pe), C.byref(ps)), "pb_reserve_j")
        assert (pe.value, ps.value) == self._ptrs, "j store moved"
        engine.check(L.pb_upload_j_range(b.epj.ctypes.data, 0, self.n_loc, C.byref(engine.LAYOUT_EPJ),
                                         b.spj.ctypes.data, 0, self.n_nodes, C.byref(engine.LAYOUT_SPJ)), "pb_upload_j_range")

    def step_device_walk(self, force, theta=0.3, resident=True):
        """One tree step with the interaction lists built on the GPU from the global tree (local +
        LET elements).  Host work: stage the tree, pack this rank's own j, name the LET rows; everything else —
        LET gather, NCCL all-to-all, tree walk, i-particle preparation, task planning, forces, reduction — runs on the
        device with no host round trip in between (resident=False: the round-1 host-planned pb_tree_force)."""
        import time
        b, L = self.batch, self.L
        t0 = time.perf_counter()
        cells, groups, em = self._stage_tree()
        engine.check(L.pb_set_params(self.prm["eps"] ** 2, self.prm["r_out"] ** 2, self.prm["G"]), "pb_set_params")
        engine.check(L.pb_tree_upload_let(cells.ctypes.data, len(cells), groups.ctypes.data, len(groups), float(theta),
                                          em.ctypes.data, len(em)), "pb_tree_upload_let")       # starts the tree walk
        t1 = time.perf_counter()
        self.upload_local_j()
        t2 = time.perf_counter()
        self.exchange_device()
        t3 = time.perf_counter()
        if resident:
            engine.check(L.pb_tree_force_resident(force.ctypes.data, C.byref(engine.LAYOUT_FORCE)), "pb_tree_force_resident")
        else:
            engine.check(L.pb_tree_force(b.epi.ctypes.data, C.byref(engine.LAYOUT_EPI), force.ctypes.data, C.byref(engine.LAYOUT_FORCE)), "pb_tree_force")
        t4 = time.perf_counter()
        for k, v in (("tree_upload", t1 - t0), ("upload_local_j", t2 - t1), ("let_exchange_enqueue", t3 - t2), ("walk_plan_force_wait", t4 - t3)):
            self.host_dw[k] = self.host_dw.get(k, 0.0) + v
        self.n_dw += 1
        return force

    def let_exchange_only(self):
        """Device gather + all-to-all + publish of one step with the local j already resident (bench `value` leg)."""
        self.exchange_device()

    def step(self, force):
        import time
        b, L = self.batch, self.L
        t0 = time.perf_counter()
        engine.check(L.pb_set_params(self.prm["eps"] ** 2, self.prm["r_out"] ** 2, self.prm["G"]), "pb_set_params")
        self.upload_local_j()
        t1 = time.perf_counter()
        self.exchange_device()
        t3 = time.perf_counter()
        if getattr(self, "_tables_for", None) is not force:
            self._tables, self._tables_for = engine.make_dispatch_tables(b, force), force
        f = engine.calc_force_all_and_write_back(b, self.prm["eps"], self.prm["r_out"], self.prm["G"],
                                                 force=force, my_rank=self.rank, send=False, tables=self._tables)
        t4 = time.perf_counter()
        for k, v in (("upload_local_j", t1 - t0), ("let_exchange_enqueue", t3 - t1), ("dispatch_retrieve_loop", t4 - t3)):
            self.host_s[k] = self.host_s.get(k, 0.0) + v
        self.n_steps += 1
        return f
