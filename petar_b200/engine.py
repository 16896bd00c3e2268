"""Python face of the native soft-force engine.

Two layers, both thin:

* :data:`lib` — the raw C ABI of ``libpetar_b200.so`` (``include/petar_b200.h``) via ctypes.
* :class:`CalcForceWithLinearCutoffCUDAMultiWalk` / :func:`RetrieveForceCUDA` — the reference's
  dispatch / retrieve functors (``src/force_gpu_cuda.hpp:103-133, 165-168``) with the same names,
  argument order and meaning, routed through the C++ shim ``force_gpu_b200.cpp`` (the object that
  replaces ``build/force_gpu_cuda.o``) so that tests drive the very symbols PeTar would link.
* :func:`calc_force_all_and_write_back` — the loop FDPS ``calcForceAllAndWriteBackMultiWalkIndex``
  runs around those functors (call site ``src/petar.hpp:894-899``) over a prebuilt
  :class:`~petar_b200.walks.WalkBatch`.

There is no CPU fallback anywhere in this module: if the native libraries are missing it raises,
and without a CUDA device every call fails with ``PB_ERR_NO_DEVICE``.
"""
import ctypes as C
import os

import numpy as np

from .types import EPISoft, EPJSoft, SPJQuad, ForceSoft, PtclCorr, LARGE_FLOAT

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(_HERE, "lib")

_vp = C.c_void_p


class PbError(RuntimeError):
    def __init__(self, code, where, msg):
        super().__init__(f"{where} failed ({code}): {msg}")
        self.code = code


class LayoutEpi(C.Structure):
    _fields_ = [("stride", C.c_size_t), ("off_pos", C.c_size_t), ("off_rsearch", C.c_size_t)]


class LayoutEpj(C.Structure):
    _fields_ = [("stride", C.c_size_t), ("off_pos", C.c_size_t), ("off_mass", C.c_size_t), ("off_rsearch", C.c_size_t)]


class LayoutSpj(C.Structure):
    _fields_ = [("stride", C.c_size_t), ("off_pos", C.c_size_t), ("off_mass", C.c_size_t), ("off_quad", C.c_size_t), ("has_quad", C.c_int)]


class LayoutForce(C.Structure):
    _fields_ = [("stride", C.c_size_t), ("off_acc", C.c_size_t), ("off_pot", C.c_size_t), ("off_nngb", C.c_size_t)]


class LayoutCorr(C.Structure):
    _fields_ = [(k, C.c_size_t) for k in ("stride", "off_pos", "off_mass", "off_r_in", "off_r_out", "off_id", "off_mass_backup",
                                          "off_status", "off_acc", "off_pot_tot", "off_pot_soft")]


class CorrParams(C.Structure):
    _fields_ = [("eps2", C.c_double), ("r_out", C.c_double), ("G", C.c_double), ("status_no_cm", C.c_double), ("replay_fp32", C.c_int)]


class Profile(C.Structure):
    _fields_ = [("t_copy", C.c_double), ("t_send", C.c_double), ("t_recv", C.c_double), ("t_calc", C.c_double),
                ("n_walk", C.c_longlong), ("n_epi", C.c_longlong), ("n_epj", C.c_longlong), ("n_spj", C.c_longlong),
                ("n_call", C.c_longlong), ("n_interaction_ep", C.c_longlong), ("n_interaction_sp", C.c_longlong),
                ("n_kernel_launch", C.c_longlong), ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong),
                ("t_plan", C.c_double), ("t_pack", C.c_double), ("t_unpack", C.c_double), ("t_enqueue", C.c_double),
                ("t_gap", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _off(dt, name):
    return dt.fields[name][1]


LAYOUT_EPI = LayoutEpi(EPISoft.itemsize, _off(EPISoft, "pos"), _off(EPISoft, "r_search"))
LAYOUT_EPJ = LayoutEpj(EPJSoft.itemsize, _off(EPJSoft, "pos"), _off(EPJSoft, "mass"), _off(EPJSoft, "r_search"))
LAYOUT_SPJ = LayoutSpj(SPJQuad.itemsize, _off(SPJQuad, "pos"), _off(SPJQuad, "mass"), _off(SPJQuad, "quad"), 1)
LAYOUT_FORCE = LayoutForce(ForceSoft.itemsize, _off(ForceSoft, "acc"), _off(ForceSoft, "pot"), _off(ForceSoft, "n_ngb"))

ABI_VERSION = 6   # PB_ABI_VERSION of include/petar_b200.h this module binds

# every symbol include/petar_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "pb_init", "pb_finalize", "pb_abi_version", "pb_last_error", "pb_set_params", "pb_set_option", "pb_get_option",
    "pb_upload_j", "pb_dispatch_index", "pb_dispatch_direct", "pb_retrieve", "pb_get_profile",
    "pb_record_begin", "pb_record_end", "pb_replay", "pb_replay_launches",
    "pb_reserve_j", "pb_upload_j_range", "pb_publish_j", "pb_pack_epj_host", "pb_pack_epj_host_indexed", "pb_pack_spj_host",
    "pb_field_at_points", "pb_dispatch_count_index", "pb_tree_upload", "pb_tree_force", "pb_tree_lists",
    "pb_correct_changeover", "pb_retrieve_neighbors", "pb_tree_upload_let", "pb_debug_plan", "pb_tree_stage",
    "pb_tree_force_resident", "pb_tree_timeline", "pb_let_gather_epj", "pb_let_pack_spj", "pb_stream_wait_upload",
]

_lib = None
_shim = None
_shim_direct = None


def _need(path):
    if not os.path.exists(path):
        raise RuntimeError(
            f"native library {path} is missing: build it with `python -m petar_b200.build` "
            "(petar_b200 has no CPU fallback)")
    return path


def load():
    """Load libpetar_b200.so and declare the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    # PETAR_B200_LIB: load another build of the same library (kernel tuning experiments, tools/)
    L = C.CDLL(_need(os.environ.get("PETAR_B200_LIB") or os.path.join(LIBDIR, "libpetar_b200.so")), mode=C.RTLD_GLOBAL)
    L.pb_init.argtypes = [C.c_int, C.c_int]
    L.pb_finalize.restype = None
    L.pb_last_error.restype = C.c_char_p
    L.pb_set_params.argtypes = [C.c_double, C.c_double, C.c_double]
    L.pb_set_option.argtypes = [C.c_char_p, C.c_longlong]
    L.pb_get_option.argtypes = [C.c_char_p, C.POINTER(C.c_longlong)]
    L.pb_upload_j.argtypes = [_vp, C.c_int, C.POINTER(LayoutEpj), _vp, C.c_int, C.POINTER(LayoutSpj)]
    L.pb_dispatch_index.argtypes = [C.c_int, _vp, _vp, C.POINTER(LayoutEpi), _vp, _vp, _vp, _vp]
    L.pb_dispatch_direct.argtypes = [C.c_int, _vp, _vp, C.POINTER(LayoutEpi), _vp, _vp, C.POINTER(LayoutEpj), _vp, _vp, C.POINTER(LayoutSpj)]
    L.pb_dispatch_count_index.argtypes = [C.c_int, _vp, _vp, C.POINTER(LayoutEpi), _vp, _vp]
    L.pb_retrieve.argtypes = [C.c_int, _vp, _vp, C.POINTER(LayoutForce)]
    L.pb_get_profile.argtypes = [C.POINTER(Profile), C.c_int]
    L.pb_replay.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.pb_reserve_j.argtypes = [C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp)]
    L.pb_upload_j_range.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(LayoutEpj), _vp, C.c_int, C.c_int, C.POINTER(LayoutSpj)]
    L.pb_publish_j.argtypes = [_vp]
    L.pb_let_gather_epj.argtypes = [_vp, C.c_int, _vp]
    L.pb_let_pack_spj.argtypes = [_vp, C.c_int, C.POINTER(LayoutSpj), _vp]
    L.pb_stream_wait_upload.argtypes = [_vp]
    L.pb_pack_epj_host.argtypes = [_vp, C.c_int, C.POINTER(LayoutEpj), _vp]
    L.pb_pack_epj_host_indexed.argtypes = [_vp, _vp, C.c_int, C.POINTER(LayoutEpj), _vp]
    L.pb_pack_spj_host.argtypes = [_vp, C.c_int, C.POINTER(LayoutSpj), _vp]
    L.pb_tree_upload.argtypes = [_vp, C.c_int, _vp, C.c_int, C.c_double]
    L.pb_tree_stage.argtypes = [C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp)]
    L.pb_tree_upload_let.argtypes = [_vp, C.c_int, _vp, C.c_int, C.c_double, _vp, C.c_int]
    L.pb_tree_force.argtypes = [_vp, C.POINTER(LayoutEpi), _vp, C.POINTER(LayoutForce)]
    L.pb_tree_lists.argtypes = [_vp, _vp, _vp, C.c_longlong, _vp, C.c_longlong]
    L.pb_tree_force_resident.argtypes = [_vp, C.POINTER(LayoutForce)]
    L.pb_tree_timeline.argtypes = [C.POINTER(C.c_float), C.c_int]
    L.pb_correct_changeover.argtypes = [C.c_int, _vp, C.POINTER(LayoutCorr), C.c_int, _vp, C.POINTER(LayoutCorr), _vp, _vp, C.POINTER(CorrParams)]
    L.pb_debug_plan.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
    L.pb_retrieve_neighbors.argtypes = [C.POINTER(C.c_longlong), _vp, _vp, C.c_longlong]
    L.pb_field_at_points.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double, _vp, _vp, _vp, _vp]
    _lib = L
    return L


def load_shim(direct=False):
    """Load the C++ shim that defines PeTar's functor symbols (index or non-index build)."""
    global _shim, _shim_direct
    load()
    if direct:
        if _shim_direct is None:
            S = C.CDLL(_need(os.path.join(LIBDIR, "libpetar_b200_shim_direct.so")))
            S.pb_shim_dispatch_direct.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                                  _vp, _vp, _vp, _vp, _vp, _vp]
            S.pb_shim_retrieve.argtypes = [C.c_int, C.c_int, _vp, _vp]
            _shim_direct = S
        return _shim_direct
    if _shim is None:
        S = C.CDLL(_need(os.path.join(LIBDIR, "libpetar_b200_shim.so")))
        S.pb_shim_dispatch.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int]
        S.pb_shim_retrieve.argtypes = [C.c_int, C.c_int, _vp, _vp]
        S.pb_shim_dispatch_count.argtypes = [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int]
        S.pb_shim_walk_group_loop.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, _vp,
                                              _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int]
        S.pb_shim_profile.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]
        S.pb_shim_profile.restype = None
        _shim = S
    return _shim


def check(rc, where):
    if rc != 0:
        raise PbError(rc, where, load().pb_last_error().decode())


def set_option(key, value):
    check(load().pb_set_option(key.encode(), int(value)), f"pb_set_option({key})")


def get_option(key):
    v = C.c_longlong(0)
    check(load().pb_get_option(key.encode(), C.byref(v)), f"pb_get_option({key})")
    return int(v.value)


def get_profile(reset=False):
    p = Profile()
    load().pb_get_profile(C.byref(p), int(reset))
    return p.as_dict()


def _ptr(a):
    return a.ctypes.data if a is not None and len(a) else None


# ---------------------------------------------------------------------------------------------
# the reference's functor interface (src/force_gpu_cuda.hpp), same names and argument order
# ---------------------------------------------------------------------------------------------
class CalcForceWithLinearCutoffCUDAMultiWalk:
    """Dispatch functor, index mode.  State is only (my_rank, eps2, rcut2, G), as in the reference;
    PeTar constructs a temporary of it every tree step (src/petar.hpp:894)."""

    def __init__(self, my_rank=0, eps2=0.0, rcut2=0.0, G=1.0):
        self.initialize(my_rank, eps2, rcut2, G)

    def initialize(self, my_rank, eps2, rcut2, G):
        self.my_rank, self.eps2, self.rcut2, self.G = int(my_rank), float(eps2), float(rcut2), float(G)

    def __call__(self, tag, n_walk, epi, n_epi, id_epj, n_epj, id_spj, n_spj, epj, n_epj_tot, spj, n_spj_tot, send_flag):
        """epi / id_epj / id_spj: uint64 arrays of per-walk pointers; n_*: int32 arrays;
        epj / spj: the shared sorted arrays (EPJSoft / SPJQuad)."""
        S = load_shim()
        assert epj.dtype == EPJSoft and spj.dtype == SPJQuad
        return S.pb_shim_dispatch(self.my_rank, self.eps2, self.rcut2, self.G, int(tag), int(n_walk),
                                  _ptr(epi), _ptr(n_epi), _ptr(id_epj), _ptr(n_epj), _ptr(id_spj), _ptr(n_spj),
                                  _ptr(epj), int(n_epj_tot), _ptr(spj), int(n_spj_tot), int(bool(send_flag)))


class CalcForceWithLinearCutoffCUDA:
    """Dispatch functor, non-index mode: per-walk j arrays (src/force_gpu_cuda.hpp:137-162)."""

    def __init__(self, my_rank=0, eps2=0.0, rcut2=0.0, G=1.0):
        self.my_rank, self.eps2, self.rcut2, self.G = int(my_rank), float(eps2), float(rcut2), float(G)

    def __call__(self, tag, n_walk, epi, n_epi, epj, n_epj, spj, n_spj):
        S = load_shim(direct=True)
        return S.pb_shim_dispatch_direct(self.my_rank, self.eps2, self.rcut2, self.G, int(tag), int(n_walk),
                                         _ptr(epi), _ptr(n_epi), _ptr(epj), _ptr(n_epj), _ptr(spj), _ptr(n_spj))


def RetrieveForceCUDA(tag, n_walk, ni, force, direct=False):
    """Retrieve functor (src/force_gpu_cuda.hpp:165-168): ASSIGNS force[iw][i].{acc,pot,n_ngb}."""
    return load_shim(direct).pb_shim_retrieve(int(tag), int(n_walk), _ptr(ni), _ptr(force))


class SearchNeighborCUDAMultiWalk:
    """EXTENSION: multiwalk dispatch functor for PeTar's neighbour-search tree
    (tree_nb), the GPU form of SearchNeighborEpEpSimd (reference src/soft_force.hpp:239-283): same call
    signature as the force functor minus the superparticle arguments."""

    def __init__(self, my_rank=0):
        self.my_rank = int(my_rank)

    def __call__(self, tag, n_walk, epi, n_epi, id_epj, n_epj, epj, n_epj_tot, send_flag):
        S = load_shim()
        return S.pb_shim_dispatch_count(self.my_rank, int(tag), int(n_walk), _ptr(epi), _ptr(n_epi), _ptr(id_epj), _ptr(n_epj),
                                        _ptr(epj), int(n_epj_tot), int(bool(send_flag)))


def debug_plan(n_epi, n_epj, n_spj, n_streams_active=1):
    """Host-only test hook: (walks[n_walk, 6], tasks[n_tasks, 10], iblocks[n_iblocks, 5], n_part) of the task
    plan for one sub-batch with these list lengths (see pb_debug_plan in include/petar_b200.h)."""
    L = load()
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (n_epi, n_epj, n_spj)]
    nw = len(a[0])
    walks = np.zeros((nw, 6), dtype=np.int32)
    nib, npart = C.c_int(0), C.c_longlong(0)
    nt = L.pb_debug_plan(nw, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, int(n_streams_active),
                         walks.ctypes.data, None, 0, None, 0, C.byref(nib), C.byref(npart))
    if nt < 0:
        check(nt, "pb_debug_plan")
    tasks = np.zeros((max(nt, 1), 10), dtype=np.int32)
    ibl = np.zeros((max(nib.value, 1), 5), dtype=np.int32)
    L.pb_debug_plan(nw, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, int(n_streams_active),
                    walks.ctypes.data, tasks.ctypes.data, nt, ibl.ctypes.data, nib.value, C.byref(nib), C.byref(npart))
    return walks, tasks[:nt], ibl[:nib.value], npart.value


def retrieve_neighbors(n_i):
    """CSR neighbour lists (nb_off[n_i + 1], nb_idx) of the last retrieved count dispatch (option nb_lists)."""
    L = load()
    n = C.c_longlong(0)
    off = np.zeros(n_i + 1, dtype=np.int32)
    check(L.pb_retrieve_neighbors(C.byref(n), off.ctypes.data, None, 0), "pb_retrieve_neighbors")
    idx = np.zeros(n.value, dtype=np.int32)
    if n.value:
        check(L.pb_retrieve_neighbors(C.byref(n), None, idx.ctypes.data, n.value), "pb_retrieve_neighbors")
    return off, idx


def tree_neighbor_search(batch, n_walk_limit=200, force=None, lists=False, tables=None):
    """What ``PeTar::treeNeighborSearch`` (reference src/petar.hpp:767-788) would do with the extension
    functor: count neighbours over every walk's EP list; only ``n_ngb`` of ``force`` is assigned.
    lists=True also returns the neighbour lists themselves, (f, nb_off, nb_idx): CSR over the i-particles in
    batch order, indices into batch.epj — what ``getNeighborListOneParticle`` returns particle by particle."""
    f = np.zeros(batch.n_epi_total, dtype=ForceSoft) if force is None else force
    none_u64, none_i32 = np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.int32)
    disp = SearchNeighborCUDAMultiWalk(0)
    if lists:
        set_option("nb_lists", 1)
    offs, idxs = [], []

    def collect(t):
        RetrieveForceCUDA(0, t.n_walk, t.n_epi, t.force_ptrs)
        if lists:
            o, i = retrieve_neighbors(int(t.n_epi.sum()))
            offs.append(o); idxs.append(i)

    try:
        assert disp(0, 0, none_u64, none_i32, none_u64, none_i32, batch.epj, len(batch.epj), True) == 0
        prev = None
        if tables is None:                               # FDPS holds these ready; pass make_dispatch_tables(batch, force) to reuse them
            tables = (batch.pointer_tables(f, slice(w0, min(w0 + n_walk_limit, batch.n_walk))) for w0 in range(0, batch.n_walk, n_walk_limit))
        for t in tables:
            if prev is not None:
                collect(prev)
            assert disp(0, t.n_walk, t.epi_ptrs, t.n_epi, t.id_epj_ptrs, t.n_epj, batch.epj, len(batch.epj), False) == 0
            prev = t
        if prev is not None:
            collect(prev)
    finally:
        if lists:
            set_option("nb_lists", 0)
    if not lists:
        return f
    nb_idx = np.concatenate(idxs) if idxs else np.zeros(0, dtype=np.int32)
    nb_off = np.zeros(batch.n_epi_total + 1, dtype=np.int64)
    k = 0
    base = 0
    for o in offs:
        n = len(o) - 1
        nb_off[k + 1:k + n + 1] = base + o[1:]
        k += n; base += int(o[-1])
    return f, nb_off.astype(np.int32), nb_idx


TIMELINE_KEYS = ("walk", "wait_j", "iprep_plan", "force", "reduce", "d2h")


def tree_timeline():
    """Device timeline (ms per phase) of the last device-resident tree step (pb_tree_timeline)."""
    ms = (C.c_float * 6)()
    check(load().pb_tree_timeline(ms, 6), "pb_tree_timeline")
    return dict(zip(TIMELINE_KEYS, (float(x) for x in ms)))


def tree_force(batch, cells, groups, eps, r_out, G, theta=0.3, force=None, upload=True, elem_map=None, resident=False):
    """Device-side list building: publish j, upload the tree, let the GPU build every
    group's id_epj / id_spj and run the force kernels on them.  `batch` only supplies epj / spj / epi
    (its host index lists are NOT used).  Returns ForceSoft[n_epi_total] in group order."""
    L = load()
    f = np.zeros(batch.n_epi_total, dtype=ForceSoft) if force is None else force
    check(L.pb_set_params(eps * eps, r_out * r_out, G), "pb_set_params")
    if upload:
        # tree first: its upload also starts the walk's counting pass, which then runs on the GPU
        # while the host packs the j-particles
        if elem_map is None:
            check(L.pb_tree_upload(cells.ctypes.data, len(cells), groups.ctypes.data, len(groups), float(theta)), "pb_tree_upload")
        else:                                            # tree over local + LET elements, store in any order
            em = np.ascontiguousarray(elem_map, dtype=np.int32)
            check(L.pb_tree_upload_let(cells.ctypes.data, len(cells), groups.ctypes.data, len(groups), float(theta),
                                       em.ctypes.data, len(em)), "pb_tree_upload_let")
        check(L.pb_upload_j(batch.epj.ctypes.data, len(batch.epj), C.byref(LAYOUT_EPJ),
                            batch.spj.ctypes.data, len(batch.spj), C.byref(LAYOUT_SPJ)), "pb_upload_j")
    if resident:          # i-particles come from the j store, plan and forces stay on the device (pb_tree_force_resident)
        check(L.pb_tree_force_resident(f.ctypes.data, C.byref(LAYOUT_FORCE)), "pb_tree_force_resident")
    else:
        check(L.pb_tree_force(batch.epi.ctypes.data, C.byref(LAYOUT_EPI), f.ctypes.data, C.byref(LAYOUT_FORCE)), "pb_tree_force")
    return f


def tree_stage(n_cells, n_groups):
    """(cells, groups) numpy views of the library's pinned tree staging buffers (pb_tree_stage): write the tree into
    them and hand them to tree_force / pb_tree_upload, which then skips its own staging copy."""
    from .types import TreeCell, TreeGroup
    pc, pg = _vp(0), _vp(0)
    check(load().pb_tree_stage(int(n_cells), int(n_groups), C.byref(pc), C.byref(pg)), "pb_tree_stage")
    cells = np.frombuffer((C.c_char * (TreeCell.itemsize * n_cells)).from_address(pc.value), dtype=TreeCell) if n_cells else np.zeros(0, TreeCell)
    groups = np.frombuffer((C.c_char * (TreeGroup.itemsize * n_groups)).from_address(pg.value), dtype=TreeGroup) if n_groups else np.zeros(0, TreeGroup)
    return cells, groups


def tree_lists(n_groups):
    """Test hook: (n_ep[g], n_sp[g], id_ep concatenated, id_sp concatenated) of the last tree_force."""
    L = load()
    ne, ns = np.zeros(n_groups, dtype=np.int32), np.zeros(n_groups, dtype=np.int32)
    check(L.pb_tree_lists(ne.ctypes.data, ns.ctypes.data, None, 0, None, 0), "pb_tree_lists")
    ide, ids = np.zeros(int(ne.sum()), dtype=np.int32), np.zeros(int(ns.sum()), dtype=np.int32)
    check(L.pb_tree_lists(ne.ctypes.data, ns.ctypes.data, ide.ctypes.data if len(ide) else None, len(ide),
                          ids.ctypes.data if len(ids) else None, len(ids)), "pb_tree_lists")
    return ne, ns, ide, ids


def layout_corr(dt, r_in="r_in", r_out="r_out", with_outputs=True):
    """pb_layout_corr of a structured dtype that has pos, mass, <r_in>, <r_out>, id, mass_backup, status
    (and acc, pot_tot, pot_soft on the i side)."""
    f = dt.fields
    out = [f[k][1] for k in ("acc", "pot_tot", "pot_soft")] if with_outputs else [0, 0, 0]
    return LayoutCorr(dt.itemsize, f["pos"][1], f["mass"][1], f[r_in][1], f[r_out][1], f["id"][1], f["mass_backup"][1], f["status"][1], *out)


def correct_force_with_cutoff_tree_neighbor(ptcl, nb_off, nb_idx, ptcl_j, eps, r_out, G, replay_fp32=False,
                                            status_no_cm=-LARGE_FLOAT):
    """The particle loop of ``SystemHard::correctForceWithCutoffTreeNeighborOMP`` (reference
    src/hard.hpp:3366-3377, one particle: :1655-1691, one pair: :1408-1476) on the device: ``ptcl``
    (structured array, e.g. types.PtclCorr) has acc / pot_tot / pot_soft corrected in place over the CSR
    neighbour lists ``nb_idx[nb_off[i]:nb_off[i+1]]`` (indices into ``ptcl_j``)."""
    nb_off = np.ascontiguousarray(nb_off, dtype=np.int32)
    nb_idx = np.ascontiguousarray(nb_idx, dtype=np.int32)
    assert len(nb_off) == len(ptcl) + 1 and ptcl.flags["C_CONTIGUOUS"] and ptcl_j.flags["C_CONTIGUOUS"]
    li, lj = layout_corr(ptcl.dtype), layout_corr(ptcl_j.dtype, with_outputs=False)
    prm = CorrParams(eps * eps, r_out, G, status_no_cm, int(bool(replay_fp32)))
    check(load().pb_correct_changeover(len(ptcl), _ptr(ptcl), C.byref(li), len(ptcl_j), _ptr(ptcl_j), C.byref(lj),
                                       _ptr(nb_off), _ptr(nb_idx), C.byref(prm)), "pb_correct_changeover")
    return ptcl


def get_gravity_and_potential_at_point(x, y, z, particles, G=1.0):
    """AMUSE's ``get_gravity_at_point`` + ``get_potential_at_point`` (reference
    amuse-interface/interface.cc:966-1030) in one call: direct sum over ``particles`` (any structured
    array with ``pos`` (3 x f8) and ``mass`` (f8) fields), eps = 0, no cutoff.  Returns (ax, ay, az, phi)."""
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    n = len(x)
    out = [np.zeros(n) for _ in range(4)]
    dt = particles.dtype
    check(load().pb_field_at_points(x.ctypes.data, y.ctypes.data, z.ctypes.data, n, particles.ctypes.data, len(particles),
                                    dt.itemsize, dt.fields["pos"][1], dt.fields["mass"][1], float(G),
                                    *(o.ctypes.data for o in out)), "pb_field_at_points")
    return tuple(out)


# ---------------------------------------------------------------------------------------------
# what FDPS does around the functors
# ---------------------------------------------------------------------------------------------
N_WALK_LIMIT = 200   # reference src/petar.hpp:888


class DispatchTables(list):
    """Per-walk-group pointer tables (list of walks.PointerTables) plus, for the C++ loop driver,
    the arrays of pointers to them."""

    def finish(self):
        def col(name):
            return np.array([getattr(t, name).ctypes.data for t in self], dtype=np.uint64)
        self.n_walk = np.array([t.n_walk for t in self], dtype=np.int32)
        self.cols = [col(k) for k in ("epi_ptrs", "n_epi", "id_epj_ptrs", "n_epj", "id_spj_ptrs", "n_spj", "force_ptrs")]
        return self


def make_dispatch_tables(batch, force, n_walk_limit=N_WALK_LIMIT):
    """The per-walk-group pointer tables FDPS holds ready when it calls dispatch (built once for a
    prebuilt WalkBatch so that the emulated FDPS loop does not pay numpy bookkeeping per group)."""
    return DispatchTables(batch.pointer_tables(force, slice(w0, min(w0 + n_walk_limit, batch.n_walk)))
                          for w0 in range(0, batch.n_walk, n_walk_limit)).finish()


def calc_force_all_and_write_back(batch, eps, r_out, G, n_walk_limit=N_WALK_LIMIT, my_rank=0, force=None, send=True, tables=None,
                                  python_loop=False):
    """Emulates ``tree_soft.calcForceAllAndWriteBackMultiWalkIndex(dispatch, retrieve, 1, ..., n_walk_limit)``
    (src/petar.hpp:888-899) over a prebuilt WalkBatch: one dispatch with send_flag=true publishing all
    j, then per walk group dispatch(send_flag=false) and — after the next group's lists would have
    been built — retrieve of the previous group.  Returns ForceSoft[n_epi_total].
    `tables` (optional): make_dispatch_tables(batch, force, n_walk_limit), reused across steps.
    The loop itself runs in C++ (pb_shim_walk_group_loop in the shim, calling the functors the way FDPS
    does); python_loop=True drives the same functors call by call from Python instead."""
    f = np.zeros(batch.n_epi_total, dtype=ForceSoft) if force is None else force
    if tables is None:
        tables = make_dispatch_tables(batch, f, n_walk_limit)
    if not python_loop:
        S = load_shim()
        c = tables.cols
        rc = S.pb_shim_walk_group_loop(int(my_rank), float(eps * eps), float(r_out * r_out), float(G), len(tables), _ptr(tables.n_walk),
                                       *(_ptr(a) for a in c), _ptr(batch.epj), len(batch.epj), _ptr(batch.spj), len(batch.spj), int(bool(send)))
        assert rc == 0
        return f
    disp = CalcForceWithLinearCutoffCUDAMultiWalk(my_rank, eps * eps, r_out * r_out, G)
    none_u64 = np.zeros(0, dtype=np.uint64)
    none_i32 = np.zeros(0, dtype=np.int32)
    if send:
        rc = disp(0, 0, none_u64, none_i32, none_u64, none_i32, none_u64, none_i32,
                  batch.epj, len(batch.epj), batch.spj, len(batch.spj), True)
        assert rc == 0
    prev = None
    for t in tables:
        if prev is not None:
            RetrieveForceCUDA(0, prev.n_walk, prev.n_epi, prev.force_ptrs)
        disp = CalcForceWithLinearCutoffCUDAMultiWalk(my_rank, eps * eps, r_out * r_out, G)
        rc = disp(0, t.n_walk, t.epi_ptrs, t.n_epi, t.id_epj_ptrs, t.n_epj, t.id_spj_ptrs, t.n_spj,
                  batch.epj, len(batch.epj), batch.spj, len(batch.spj), False)
        assert rc == 0
        prev = t
    if prev is not None:
        RetrieveForceCUDA(0, prev.n_walk, prev.n_epi, prev.force_ptrs)
    return f
