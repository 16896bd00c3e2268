"""Synthetic-input harness: Plummer models, PeTar's automatic parameters, and FDPS-like walk lists.

Everything here produces INPUTS for the hot path (the part FDPS and PeTar's driver own in a real
run); none of it is on the measured path.  Native part: ``harness/tree_walk.cpp``.
"""
import ctypes as C
import math
import os

import numpy as np

from .types import EPISoft, EPJSoft, SPJQuad
from .walks import WalkBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_vp = C.c_void_p
_h = None

THETA = 0.3            # reference src/petar.hpp:153  (-T)
N_LEAF_LIMIT = 20      # reference src/petar.hpp:154
N_GROUP_LIMIT = 512    # reference src/petar.hpp:155-159


def lib():
    global _h
    if _h is None:
        path = os.path.join(_HERE, "lib", "libpetar_b200_harness.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -m petar_b200.build`")
        L = C.CDLL(path)
        L.hz_build.restype = _vp
        L.hz_build.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_double, C.c_int, C.c_int]
        L.hz_sizes.argtypes = [_vp, _vp]
        L.hz_export.argtypes = [_vp] * 9
        L.hz_export_let_sp_src.argtypes = [_vp, _vp]
        L.hz_local_boxes.argtypes = [_vp, _vp]
        L.hz_make_let.argtypes = [_vp, _vp, C.c_double, _vp, _vp, _vp, _vp]
        L.hz_free.argtypes = [_vp]
        L.hz_make_plummer.argtypes = [C.c_double, C.c_longlong, C.c_longlong, _vp, _vp, _vp, C.c_double, C.c_uint]
        _h = L
    return _h


# ---------------------------------------------------------------------------------------------
# particles and parameters
# ---------------------------------------------------------------------------------------------
def make_plummer(n, mass_glb=1.0, eng=-0.25, rank_seed=0):
    """Equal-mass Plummer model in Henon units (reference src/particle_distribution_generator.hpp:173-250)."""
    mass = np.empty(n)
    pos = np.empty((n, 3))
    vel = np.empty((n, 3))
    lib().hz_make_plummer(mass_glb, n, n, mass.ctypes.data, pos.ctypes.data, vel.ctypes.data, eng, rank_seed)
    return mass, pos, vel


def regular_time_step(dt):
    """reference src/petar.hpp:755-764"""
    r = 1.0
    if dt < 1:
        while r > dt:
            r *= 0.5
    else:
        while r <= dt:
            r *= 2.0
        r *= 0.5
    return r


def petar_auto_params(mass, vel, G=1.0, ratio_r_cut=0.1, search_vel_factor=3.0):
    """PeTar's automatic r_out / r_in / dt_soft / r_search_min (reference src/petar.hpp:3156-3239),
    singles only."""
    n = len(mass)
    m_tot = mass.sum()
    vcm = (mass[:, None] * vel).sum(0) / m_tot
    dv = vel - vcm
    vel_disp = math.sqrt((dv * dv).sum() / 3.0 / n)
    r_out = 0.1 * G * m_tot / (n ** (1.0 / 3.0)) / (3 * vel_disp * vel_disp)
    r_in = r_out * ratio_r_cut
    dt_soft = regular_time_step(0.1 * r_out / vel_disp)
    r_search_min = search_vel_factor * vel_disp * dt_soft + r_out
    return dict(r_out=r_out, r_in=r_in, dt_soft=dt_soft, r_search_min=r_search_min, vel_disp=vel_disp,
                search_factor=search_vel_factor, mean_mass=m_tot / n, G=G, eps=0.0)


def particle_rout_rsearch(mass, vel, prm):
    """Per-particle changeover r_out (reference src/changeover.hpp:44-52 with m_fac = m / <m>,
    src/petar.hpp:3297) and r_search (reference src/ptcl.hpp:227-231)."""
    m_fac3 = np.maximum(np.cbrt(mass / prm["mean_mass"]), 1.0)
    r_out_i = m_fac3 * prm["r_out"]
    r_in_i = m_fac3 * prm["r_in"]
    v = np.sqrt((vel * vel).sum(1))
    rs = np.maximum(v * prm["dt_soft"] * prm["search_factor"] + r_out_i, prm["r_search_min"])
    return r_in_i, r_out_i, rs


def kroupa_masses(n, rng, m_lo=0.08, m_hi=150.0):
    """Kroupa (2001) two-segment IMF above 0.08 Msun (alpha = 1.3 below 0.5, 2.3 above), by inverse
    transform; returned normalised to sum 1 (Henon units)."""
    a1, a2, mb = 1.3, 2.3, 0.5

    def seg(a, lo, hi):
        return (hi ** (1 - a) - lo ** (1 - a)) / (1 - a)

    w1 = seg(a1, m_lo, mb)
    w2 = seg(a2, mb, m_hi) * mb ** (a2 - a1)
    u = rng.random(n)
    first = u < w1 / (w1 + w2)
    m = np.empty(n)
    u1 = rng.random(first.sum())
    m[first] = (m_lo ** (1 - a1) + u1 * (mb ** (1 - a1) - m_lo ** (1 - a1))) ** (1 / (1 - a1))
    u2 = rng.random((~first).sum())
    m[~first] = (mb ** (1 - a2) + u2 * (m_hi ** (1 - a2) - mb ** (1 - a2))) ** (1 / (1 - a2))
    return m / m.sum()


# ---------------------------------------------------------------------------------------------
# trees and walk lists
# ---------------------------------------------------------------------------------------------
class TreeHandle:
    """Owns one native build (local tree [+ global tree with LET], groups, lists)."""

    def __init__(self, pos, mass, rsearch, let=None, theta=THETA, n_leaf_limit=N_LEAF_LIMIT, n_group_limit=N_GROUP_LIMIT):
        self.pos = np.ascontiguousarray(pos, dtype=np.float64)
        self.mass = np.ascontiguousarray(mass, dtype=np.float64)
        self.rs = np.ascontiguousarray(rsearch, dtype=np.float64)
        self.theta = theta
        n = len(self.mass)
        if let is None:
            lp = lm = lr = None
            nle = nls = 0
            lsp = None
        else:
            lp = np.ascontiguousarray(let["pos"], dtype=np.float64)
            lm = np.ascontiguousarray(let["mass"], dtype=np.float64)
            lr = np.ascontiguousarray(let["rsearch"], dtype=np.float64)
            lsp = np.ascontiguousarray(let["spj"], dtype=SPJQuad)
            nle, nls = len(lm), len(lsp)
        self._let = (lp, lm, lr, lsp)
        p = lambda a: a.ctypes.data if a is not None and len(a) else None
        self.h = lib().hz_build(n, p(self.pos), p(self.mass), p(self.rs), nle, p(lp), p(lm), p(lr), nls, p(lsp),
                                theta, n_leaf_limit, n_group_limit)
        sz = np.zeros(8, dtype=np.int64)
        lib().hz_sizes(self.h, sz.ctypes.data)
        self.n_epj, self.n_spj, self.n_walk, self.n_epi, self.n_id_epj, self.n_id_spj, self.n_nodes, self.n_let_sp = (int(x) for x in sz)

    def export(self):
        epj_src = np.zeros(self.n_epj, dtype=np.int32)
        epi_src = np.zeros(self.n_epi, dtype=np.int32)
        spj = np.zeros(self.n_spj, dtype=SPJQuad)
        i_off = np.zeros(self.n_walk + 1, dtype=np.int64)
        ej_off = np.zeros(self.n_walk + 1, dtype=np.int64)
        sj_off = np.zeros(self.n_walk + 1, dtype=np.int64)
        id_epj = np.zeros(self.n_id_epj, dtype=np.int32)
        id_spj = np.zeros(self.n_id_spj, dtype=np.int32)
        lib().hz_export(self.h, epj_src.ctypes.data, epi_src.ctypes.data, spj.ctypes.data, i_off.ctypes.data,
                        ej_off.ctypes.data, sj_off.ctypes.data, id_epj.ctypes.data, id_spj.ctypes.data)
        return epj_src, epi_src, spj, i_off, ej_off, sj_off, id_epj, id_spj

    def let_sp_src(self):
        """for spj[n_nodes + k]: index of that entry in the `let["spj"]` array given at build time"""
        out = np.zeros(self.n_let_sp, dtype=np.int32)
        if self.n_let_sp:
            lib().hz_export_let_sp_src(self.h, out.ctypes.data)
        return out

    def local_boxes(self):
        out = np.zeros(12)
        lib().hz_local_boxes(self.h, out.ctypes.data)
        return out

    def make_let(self, remote_boxes):
        """(local particle indices to send as EP, SPJQuad array to send as SP) for one remote domain."""
        rb = np.ascontiguousarray(remote_boxes, dtype=np.float64)
        ne, ns = C.c_longlong(0), C.c_longlong(0)
        lib().hz_make_let(self.h, rb.ctypes.data, self.theta, C.byref(ne), C.byref(ns), None, None)
        ep = np.zeros(ne.value, dtype=np.int32)
        sp = np.zeros(ns.value, dtype=SPJQuad)
        lib().hz_make_let(self.h, rb.ctypes.data, self.theta, C.byref(ne), C.byref(ns),
                          ep.ctypes.data if len(ep) else None, sp.ctypes.data if len(sp) else None)
        return ep, sp

    def close(self):
        if self.h:
            lib().hz_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_walk_batch(pos, mass, rsearch, vel=None, r_in=None, r_out=None, ids=None, ptype=None, let=None,
                     theta=THETA, n_leaf_limit=N_LEAF_LIMIT, n_group_limit=N_GROUP_LIMIT, rank=0):
    """Tree + walks for one domain -> :class:`WalkBatch` in FDPS's dispatch shape.

    `let` (optional): dict(pos, mass, rsearch, spj) of entries received from other domains.
    Returns (batch, epi_src) where epi_src[k] is the local particle index of the k-th i-particle."""
    t = TreeHandle(pos, mass, rsearch, let, theta, n_leaf_limit, n_group_limit)
    epj_src, epi_src, spj, i_off, ej_off, sj_off, id_epj, id_spj = t.export()
    n_loc = len(mass)
    if let is not None and len(let["mass"]):
        apos = np.concatenate([np.asarray(pos), np.asarray(let["pos"])])
        amass = np.concatenate([np.asarray(mass), np.asarray(let["mass"])])
        ars = np.concatenate([np.asarray(rsearch), np.asarray(let["rsearch"])])
    else:
        apos, amass, ars = np.asarray(pos), np.asarray(mass), np.asarray(rsearch)
    epj = np.zeros(t.n_epj, dtype=EPJSoft)
    epj["id"] = epj_src + 1 if ids is None else np.where(epj_src < n_loc, np.asarray(ids)[np.minimum(epj_src, n_loc - 1)], epj_src + 1)
    epj["mass"] = amass[epj_src]
    epj["pos"] = apos[epj_src]
    epj["r_search"] = ars[epj_src]
    loc = epj_src < n_loc
    if vel is not None:
        epj["vel"][loc] = np.asarray(vel)[epj_src[loc]]
    if r_in is not None:
        epj["r_in"][loc] = np.asarray(r_in)[epj_src[loc]]
    if r_out is not None:
        epj["r_out"][loc] = np.asarray(r_out)[epj_src[loc]]
    epj["r_scale_next"] = 1.0
    epj["rank_org"] = rank
    epj["adr_org"] = epj_src
    epi = np.zeros(t.n_epi, dtype=EPISoft)
    epi["id"] = epi_src + 1 if ids is None else np.asarray(ids)[epi_src]
    epi["pos"] = np.asarray(pos)[epi_src]
    epi["r_search"] = np.asarray(rsearch)[epi_src]
    epi["rank_org"] = rank
    epi["type"] = 1 if ptype is None else np.asarray(ptype)[epi_src]
    batch = WalkBatch(epj, spj, epi, i_off, id_epj, ej_off, id_spj, sj_off)
    batch.tree = t
    return batch, epi_src


def plummer_case(n, rank_seed=0, theta=THETA, n_group_limit=N_GROUP_LIMIT, n_leaf_limit=N_LEAF_LIMIT):
    """Config 2/3 skeleton: equal-mass Plummer + PeTar's automatic parameters + walk lists."""
    mass, pos, vel = make_plummer(n, rank_seed=rank_seed)
    prm = petar_auto_params(mass, vel)
    r_in, r_out, rs = particle_rout_rsearch(mass, vel, prm)
    batch, epi_src = build_walk_batch(pos, mass, rs, vel=vel, r_in=r_in, r_out=r_out, theta=theta,
                                      n_group_limit=n_group_limit, n_leaf_limit=n_leaf_limit)
    return batch, epi_src, prm, (mass, pos, vel)
