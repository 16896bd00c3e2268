"""ctypes binding of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / ``--impl reference``)
may import this module.  The product package ``petar_b200`` never does.

Two libraries:

* ``oracle/lib/liboracle_soft_force.so`` — fp64 restatement of the reference's NoSimd functors
  (``oracle_soft_force.c``; reference ``src/soft_force.hpp:11-34, 38-87, 125-158, 160-200``).
* ``oracle/_ref/libpetar_ref_{avx2,avx512}.so`` — the reference's own ``PhantomGrapeQuad`` kernels
  compiled from ``/root/reference/src/phantomquad_for_p3t_x86.hpp`` (``ref_simd.cpp``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from petar_b200.types import EPISoft, EPJSoft, SPJQuad, ForceSoft, PtclCorr, LARGE_FLOAT

_HERE = os.path.dirname(os.path.abspath(__file__))
_vp = C.c_void_p
_ip = C.POINTER(C.c_int)


def build(verbose=False):
    """Compile the oracle (always) and oracle/_ref (only where /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":")[1].split())
                    return {"avx512f", "avx512dq"} <= fl
    except OSError:
        pass
    return False


_oracle = None
_ref = {}


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(_HERE, "lib", "liboracle_soft_force.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_search_neighbor_epep.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp]
        L.orc_force_epep_linear_cutoff.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, C.c_double, C.c_double]
        L.orc_force_epsp_mono.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, C.c_double]
        L.orc_force_epsp_quad.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, C.c_double]
        L.orc_force_pp.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double]
        L.orc_walks_index.argtypes = [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double]
        L.orc_make_plummer.argtypes = [C.c_double, C.c_longlong, C.c_longlong, _vp, _vp, _vp, C.c_double, C.c_uint32]
        L.orc_simdtest_inputs.argtypes = [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]
        L.orc_srand.argtypes = [C.c_uint]
        L.orc_mt_init.argtypes = [_vp, C.c_uint32]
        L.orc_mt_int32.argtypes = [_vp]
        L.orc_mt_int32.restype = C.c_uint32
        L.orc_mt_res53.argtypes = [_vp]
        L.orc_mt_res53.restype = C.c_double
        L.orc_mt_real2.argtypes = [_vp]
        L.orc_mt_real2.restype = C.c_double
        L.orc_calc_rsearch.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_calc_rsearch.restype = C.c_double
        L.orc_changeover_w.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_changeover_pair.argtypes = [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_correct_force_tree_neighbor.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        _oracle = L
    return _oracle


def ref_available(isa=None):
    isa = isa or ("avx512" if _cpu_has_avx512() else "avx2")
    return os.path.exists(os.path.join(_HERE, "_ref", f"libpetar_ref_{isa}.so"))


def ref_lib(isa=None):
    """The reference's own SIMD kernels; isa in {'avx2','avx512'} (default: best the CPU has)."""
    isa = isa or ("avx512" if _cpu_has_avx512() else "avx2")
    if isa == "avx512" and not _cpu_has_avx512():
        raise RuntimeError("host CPU lacks AVX-512")
    if isa not in _ref:
        path = os.path.join(_HERE, "_ref", f"libpetar_ref_{isa}.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.ref_simd_isa.restype = C.c_char_p
        L.ref_epep_simd.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, C.c_double, C.c_double]
        L.ref_epsp_quad_simd.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, C.c_double]
        L.ref_nb_simd.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp]
        L.ref_walks_index.argtypes = [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_int]
        L.ref_walks_index.restype = C.c_double
        _ref[isa] = L
    return _ref[isa]


_nosimd = None


def ref_nosimd_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libpetar_ref_nosimd.so"))


def ref_nosimd_lib():
    """The reference's own fp64 NoSimd functors (src/soft_force.hpp:10-236) compiled from the reference source
    (ref_nosimd.cpp): the bit-for-bit pin of the restatement in oracle_soft_force.c."""
    global _nosimd
    if _nosimd is None:
        path = os.path.join(_HERE, "_ref", "libpetar_ref_nosimd.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.ref_nosimd_search.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp]
        L.ref_nosimd_epep.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, C.c_double, C.c_double]
        L.ref_nosimd_epsp_mono.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, C.c_double]
        L.ref_nosimd_epsp_quad.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double, C.c_double]
        L.ref_nosimd_pp.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_double]
        _nosimd = L
    return _nosimd


def ref_nosimd(kind, epi, j, eps=0.0, r_out=0.0, G=1.0, force=None):
    """kind in {'search', 'epep', 'epsp_mono', 'epsp_quad', 'pp'}: the reference functor of that name on flat arrays."""
    f = new_force(len(epi)) if force is None else force
    L = ref_nosimd_lib()
    jt = SPJQuad if kind.startswith("epsp") else EPJSoft
    a = (_chk(epi, EPISoft), len(epi), _chk(j, jt), len(j), _chk(f, ForceSoft))
    if kind == "search":
        L.ref_nosimd_search(*a)
    elif kind == "epep":
        L.ref_nosimd_epep(*a, eps, r_out, G)
    elif kind == "epsp_mono":
        L.ref_nosimd_epsp_mono(*a, eps, G)
    elif kind == "epsp_quad":
        L.ref_nosimd_epsp_quad(*a, eps, G)
    elif kind == "pp":
        L.ref_nosimd_pp(*a, G)
    else:
        raise ValueError(kind)
    return f


def _chk(a, dt):
    assert a.dtype == dt and a.flags["C_CONTIGUOUS"], (a.dtype, dt)
    return a.ctypes.data


def new_force(n):
    return np.zeros(n, dtype=ForceSoft)


# ---- functor-level wrappers (fp64 oracle) -------------------------------------------------
def force_epep(epi, epj, eps, r_out, G, force=None):
    f = new_force(len(epi)) if force is None else force
    oracle_lib().orc_force_epep_linear_cutoff(_chk(epi, EPISoft), len(epi), _chk(epj, EPJSoft), len(epj), _chk(f, ForceSoft), eps, r_out, G)
    return f


def force_epsp_quad(epi, spj, eps, G, force=None):
    f = new_force(len(epi)) if force is None else force
    oracle_lib().orc_force_epsp_quad(_chk(epi, EPISoft), len(epi), _chk(spj, SPJQuad), len(spj), _chk(f, ForceSoft), eps, G)
    return f


def force_epsp_mono(epi, spj, eps, G, force=None):
    f = new_force(len(epi)) if force is None else force
    oracle_lib().orc_force_epsp_mono(_chk(epi, EPISoft), len(epi), _chk(spj, SPJQuad), len(spj), _chk(f, ForceSoft), eps, G)
    return f


def force_pp(epi, epj, G, force=None):
    f = new_force(len(epi)) if force is None else force
    oracle_lib().orc_force_pp(_chk(epi, EPISoft), len(epi), _chk(epj, EPJSoft), len(epj), _chk(f, ForceSoft), G)
    return f


def search_neighbor(epi, epj, force=None):
    f = new_force(len(epi)) if force is None else force
    oracle_lib().orc_search_neighbor_epep(_chk(epi, EPISoft), len(epi), _chk(epj, EPJSoft), len(epj), _chk(f, ForceSoft))
    return f


# ---- functor-level wrappers (reference SIMD kernels) --------------------------------------
def ref_force_epep(epi, epj, eps, r_out, G, isa=None, force=None):
    f = new_force(len(epi)) if force is None else force
    ref_lib(isa).ref_epep_simd(_chk(epi, EPISoft), len(epi), _chk(epj, EPJSoft), len(epj), _chk(f, ForceSoft), eps, r_out, G)
    return f


def ref_force_epsp_quad(epi, spj, eps, G, isa=None, force=None):
    f = new_force(len(epi)) if force is None else force
    ref_lib(isa).ref_epsp_quad_simd(_chk(epi, EPISoft), len(epi), _chk(spj, SPJQuad), len(spj), _chk(f, ForceSoft), eps, G)
    return f


def ref_search_neighbor(epi, epj, isa=None, force=None):
    f = new_force(len(epi)) if force is None else force
    ref_lib(isa).ref_nb_simd(_chk(epi, EPISoft), len(epi), _chk(epj, EPJSoft), len(epj), _chk(f, ForceSoft))
    return f


# ---- multiwalk batches ---------------------------------------------------------------------
def _walk_args(batch, force):
    """batch: petar_b200.walks.WalkBatch; returns the pointer tables FDPS would pass."""
    p = batch.pointer_tables(force)
    return p


def walks_index(batch, eps, r_out, G, walk_slice=None):
    """fp64 oracle over a WalkBatch (optionally a slice of walks); returns ForceSoft[sum n_epi]."""
    force = new_force(batch.n_epi_total)
    t = batch.pointer_tables(force, walk_slice)
    oracle_lib().orc_walks_index(
        t.n_walk, t.epi_ptrs.ctypes.data, t.n_epi.ctypes.data, t.id_epj_ptrs.ctypes.data, t.n_epj.ctypes.data,
        t.id_spj_ptrs.ctypes.data, t.n_spj.ctypes.data, _chk(batch.epj, EPJSoft), _chk(batch.spj, SPJQuad),
        t.force_ptrs.ctypes.data, eps, r_out, G)
    return force


def ref_walks_index(batch, eps, r_out, G, isa=None, n_threads=0, walk_slice=None, force=None):
    """Reference SIMD CPU path over a WalkBatch; returns (ForceSoft[...], seconds)."""
    force = new_force(batch.n_epi_total) if force is None else force
    t = batch.pointer_tables(force, walk_slice)
    sec = ref_lib(isa).ref_walks_index(
        t.n_walk, t.epi_ptrs.ctypes.data, t.n_epi.ctypes.data, t.id_epj_ptrs.ctypes.data, t.n_epj.ctypes.data,
        t.id_spj_ptrs.ctypes.data, t.n_spj.ctypes.data, _chk(batch.epj, EPJSoft), _chk(batch.spj, SPJQuad),
        t.force_ptrs.ctypes.data, eps, r_out, G, int(n_threads))
    return force, sec


# ---- synthetic inputs ----------------------------------------------------------------------
def make_plummer(n, mass_glb=1.0, eng=-0.25, rank_seed=0):
    """reference makePlummerModel (src/particle_distribution_generator.hpp:173-250)."""
    mass = np.empty(n)
    pos = np.empty((n, 3))
    vel = np.empty((n, 3))
    oracle_lib().orc_make_plummer(mass_glb, n, n, mass.ctypes.data, pos.ctypes.data, vel.ctypes.data, eng, rank_seed)
    return mass, pos, vel


def simdtest_inputs(n_epi=1000, n_epj=2000, n_spj=1000):
    """The input set of reference src/simd_test.cxx (r_out=0.01, eps=1e-4, G=1)."""
    epi = np.zeros(n_epi, dtype=EPISoft)
    epj = np.zeros(n_epj, dtype=EPJSoft)
    spj = np.zeros(n_spj, dtype=SPJQuad)
    L = oracle_lib()
    L.orc_srand(1)  # glibc default seed, as an un-seeded rand() in the reference test
    L.orc_simdtest_inputs(n_epi, n_epj, n_spj, epi.ctypes.data, epj.ctypes.data, spj.ctypes.data)
    return epi, epj, spj


SIMDTEST_PARAMS = dict(eps=1e-4, r_out=0.01, G=1.0)  # reference src/simd_test.cxx:65-67


# ---- informational second baseline: the reference's own CUDA kernels on this GPU ------------
_ref_cuda = None


def ref_cuda_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libpetar_ref_cuda.so"))


def ref_cuda_step(batch, eps, r_out, G, n_walk_limit=200):
    """One tree step through the reference's CUDA kernels (device code extracted at build time from
    /root/reference/src/force_gpu_cuda.cu, host side restated in oracle/ref_cuda_driver.cu).
    Returns (ForceSoft[...], ms of the two kernels summed over dispatches, ms wall-clock of the step)."""
    global _ref_cuda
    if _ref_cuda is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libpetar_ref_cuda.so"))
        L.refcuda_step.argtypes = [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp,
                                   C.c_double, C.c_double, C.c_double, C.c_int, _vp]
        _ref_cuda = L
    force = new_force(batch.n_epi_total)
    t = batch.pointer_tables(force)
    ms = np.zeros(2, dtype=np.float32)
    rc = _ref_cuda.refcuda_step(t.n_walk, t.epi_ptrs.ctypes.data, t.n_epi.ctypes.data, t.id_epj_ptrs.ctypes.data, t.n_epj.ctypes.data,
                                t.id_spj_ptrs.ctypes.data, t.n_spj.ctypes.data, _chk(batch.epj, EPJSoft), len(batch.epj),
                                _chk(batch.spj, SPJQuad), len(batch.spj), t.force_ptrs.ctypes.data,
                                eps * eps, r_out * r_out, G, int(n_walk_limit), ms.ctypes.data)
    if rc != 0:
        raise RuntimeError("reference CUDA baseline failed")
    return force, float(ms[0]), float(ms[1])


# ---- changeover correction (oracle_changeover.c; reference src/hard.hpp:1408-1476, 1655-1691) ----
def changeover_w(r_in, r_out, dr):
    a, p = C.c_double(), C.c_double()
    oracle_lib().orc_changeover_w(r_in, r_out, dr, C.byref(a), C.byref(p))
    return a.value, p.value


def changeover_pair(pi, pj, eps, r_out, G, replay_fp32):
    """pi, pj: 1-element PtclCorr arrays; pi is updated in place."""
    oracle_lib().orc_changeover_pair(_chk(pi, PtclCorr), _chk(pj, PtclCorr), eps * eps, r_out, G, int(replay_fp32))


def correct_force_tree_neighbor(p, nb_off, nb_idx, pj, eps, r_out, G, replay_fp32, status_no_cm=-LARGE_FLOAT):
    """In-place correction of p[*].acc / pot_tot / pot_soft over CSR neighbour lists (indices into pj)."""
    assert nb_off.dtype == np.int32 and nb_idx.dtype == np.int32 and len(nb_off) == len(p) + 1
    oracle_lib().orc_correct_force_tree_neighbor(_chk(p, PtclCorr), len(p), nb_off.ctypes.data, nb_idx.ctypes.data if len(nb_idx) else None,
                                                 _chk(pj, PtclCorr), eps * eps, r_out, G, status_no_cm, int(replay_fp32))
    return p


_ref_co = {}


def ref_changeover_available():
    return all(os.path.exists(os.path.join(_HERE, "_ref", f"libpetar_ref_changeover_{k}.so")) for k in ("f32", "f64"))


def ref_changeover_lib(replay_fp32):
    """The reference's own ChangeOver class and calcAccPotShortWithLinearCutoff (ref_changeover.cpp),
    built with -DUSE_GPU (float replay) or -DP3T_64BIT (all double)."""
    key = "f32" if replay_fp32 else "f64"
    if key not in _ref_co:
        L = C.CDLL(os.path.join(_HERE, "_ref", f"libpetar_ref_changeover_{key}.so"))
        L.ref_changeover_w.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_changeover_pair.argtypes = [_vp, _vp, C.c_double, C.c_double, C.c_double]
        _ref_co[key] = L
    return _ref_co[key]


def ref_changeover_w(r_in, r_out, dr):
    a, p = C.c_double(), C.c_double()
    ref_changeover_lib(False).ref_changeover_w(r_in, r_out, dr, C.byref(a), C.byref(p))
    return a.value, p.value


def ref_changeover_pair(pi, pj, eps, r_out, G, replay_fp32):
    ref_changeover_lib(replay_fp32).ref_changeover_pair(_chk(pi, PtclCorr), _chk(pj, PtclCorr), eps, r_out, G)
