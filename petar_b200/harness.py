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
        L.hz_export_tree.argtypes = [_vp, _vp, _vp]
        L.hz_export_elem_map.argtypes = [_vp, _vp]
        L.hz_timing.argtypes = [_vp, _vp]
        L.hz_local_boxes.argtypes = [_vp, _vp]
        L.hz_make_let.argtypes = [_vp, _vp, C.c_double, _vp, _vp, _vp, _vp]
        L.hz_free.argtypes = [_vp]
        L.hz_set_skip_walk.argtypes = [C.c_int]
        L.hz_walk_group.argtypes = [_vp, C.c_int, C.c_double, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _vp, _vp]
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

    def export_tree(self, out=None):
        """(cells, groups) in the layout of pb_tree_cell / pb_tree_group, for the device-side list
        builder.  A tree with LET elements also needs :meth:`export_elem_map`.  `out` = (cells, groups)
        arrays to write into (e.g. engine.tree_stage's pinned buffers)."""
        from .types import TreeCell, TreeGroup
        cells, groups = out if out is not None else (np.zeros(self.n_nodes, dtype=TreeCell), np.zeros(self.n_walk, dtype=TreeGroup))
        assert len(cells) == self.n_nodes and len(groups) == self.n_walk and cells.dtype == TreeCell and groups.dtype == TreeGroup
        lib().hz_export_tree(self.h, cells.ctypes.data, groups.ctypes.data)
        return cells, groups

    def export_elem_map(self):
        """For the k-th Morton-sorted element of the (global) tree: its index in epj_sorted (>= 0) or
        ~(its index in the LET part of spj) (< 0) — pb_tree_upload_let's `elem_map`."""
        out = np.zeros(self.n_epj + self.n_let_sp, dtype=np.int32)
        lib().hz_export_elem_map(self.h, out.ctypes.data)
        return out

    def walk_group(self, g):
        """(id_epj, id_spj) of group g, walked on the host now (see :func:`skip_walks`): indices into epj_sorted / spj."""
        ne, ns = C.c_longlong(0), C.c_longlong(0)
        lib().hz_walk_group(self.h, int(g), self.theta, C.byref(ne), C.byref(ns), None, None)
        e, s = np.zeros(ne.value, dtype=np.int32), np.zeros(ns.value, dtype=np.int32)
        lib().hz_walk_group(self.h, int(g), self.theta, C.byref(ne), C.byref(ns), e.ctypes.data if len(e) else None, s.ctypes.data if len(s) else None)
        return e, s

    def timing(self):
        """(seconds building the tree, seconds walking it for all groups) on the host, OpenMP."""
        out = np.zeros(2)
        lib().hz_timing(self.h, out.ctypes.data)
        return float(out[0]), float(out[1])

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


def skip_walks(on):
    """on=True: TreeHandle builds trees and i-groups only, every group's lists stay empty (inputs of the device-side
    walk at sizes where host lists would not fit, e.g. BASELINE config 5); TreeHandle.walk_group spot-checks then."""
    lib().hz_set_skip_walk(int(bool(on)))


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


# ---------------------------------------------------------------------------------------------
# BASELINE config 3 stand-in: Kroupa masses, primordial binaries, artificial particles
# ---------------------------------------------------------------------------------------------
def kroupa_binary_particles(n_star, f_bin=0.1, seed=1):
    """The particle set PeTar's SOFT tree sees for an N-star Plummer cluster with a Kroupa IMF and a
    fraction f_bin of the stars in primordial binaries, every binary carrying its artificial
    particles (worst case of SURVEY §2 #16): per binary 2 members with their soft mass zeroed,
    8 zero-mass tidal-tensor probes, 1 zero-mass centre-of-mass particle and 3 orbit-sample
    pseudo-particles that carry the binary mass and are `type 0` i-particles
    (reference src/artificial_particles.hpp:427-429, src/tidal_tensor.hpp:435-441,
    src/pseudoparticle_multipole.hpp:58-60, src/soft_ptcl.hpp:282-284).  The geometry of the
    artificial particles is schematic (probes on a cube of half-size 2a, samples on the relative
    orbit): what matters for the hot path is their number, masses, types and clustering.

    Returns dict(pos, mass, vel, rs, r_in, r_out, ptype, prm, n_star, n_bin)."""
    rng = np.random.default_rng(seed)
    _, pos, vel = make_plummer(n_star)                   # centre-of-mass phase space, MT19937 seed 0
    m_star = kroupa_masses(n_star, rng)
    n_bin = int(round(0.5 * f_bin * n_star))
    prm = petar_auto_params(m_star, vel)                 # PeTar measures these on the stars
    prm["mean_mass"] = 1.0 / n_star

    # binaries = star pairs (2k, 2k+1), k < n_bin, placed around the position of star 2k
    a = np.exp(rng.uniform(np.log(1e-6), np.log(0.8 * prm["r_in"]), n_bin))       # inside r_bin = 0.8 r_in
    d = rng.normal(size=(n_bin, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    m1, m2 = m_star[0:2 * n_bin:2], m_star[1:2 * n_bin:2]
    mb = m1 + m2
    cm_pos, cm_vel = pos[0:2 * n_bin:2].copy(), vel[0:2 * n_bin:2].copy()
    p1 = cm_pos + d * (a * m2 / mb)[:, None]
    p2 = cm_pos - d * (a * m1 / mb)[:, None]
    # an orthonormal frame of the orbital plane for the orbit samples
    e2 = np.cross(d, rng.normal(size=(n_bin, 3)))
    e2 /= np.linalg.norm(e2, axis=1)[:, None]
    ph = rng.uniform(0, 2 * np.pi, n_bin)
    samples = [cm_pos + 0.5 * a[:, None] * (np.cos(ph + k * 2 * np.pi / 3)[:, None] * d + np.sin(ph + k * 2 * np.pi / 3)[:, None] * e2)
               for k in range(3)]
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64)
    probes = (cm_pos[:, None, :] + 2.0 * a[:, None, None] * corners[None, :, :]).reshape(-1, 3)

    singles = np.arange(2 * n_bin, n_star)
    pos_all = np.concatenate([p1, p2, probes, cm_pos] + samples + [pos[singles]])
    zeros = np.zeros(n_bin)
    mass_all = np.concatenate([zeros, zeros, np.zeros(8 * n_bin), zeros, mb / 3, mb / 3, mb / 3, m_star[singles]])
    vel_all = np.concatenate([cm_vel, cm_vel, np.repeat(cm_vel, 8, axis=0), cm_vel, cm_vel, cm_vel, cm_vel, vel[singles]])
    # changeover radius / r_search follow the mass the particle stands for (member: its own, artificial: the binary's)
    m_for_r = np.concatenate([m1, m2, np.repeat(mb, 8), mb, mb, mb, mb, m_star[singles]])
    r_in, r_out, rs = particle_rout_rsearch(m_for_r, vel_all, prm)
    ptype = np.ones(len(mass_all), dtype=np.int32)
    ptype[11 * n_bin:14 * n_bin] = 0                     # orbit samples: artificial, not c.m., mass > 0
    return dict(pos=pos_all, mass=mass_all, vel=vel_all, rs=rs, r_in=r_in, r_out=r_out, ptype=ptype, prm=prm,
                n_star=n_star, n_bin=n_bin, m_members=(m1, m2))


def kroupa_binary_case(n_star, f_bin=0.1, seed=1, theta=THETA, n_group_limit=N_GROUP_LIMIT, n_leaf_limit=N_LEAF_LIMIT):
    """Walk lists for :func:`kroupa_binary_particles` (single domain)."""
    P = kroupa_binary_particles(n_star, f_bin, seed)
    batch, epi_src = build_walk_batch(P["pos"], P["mass"], P["rs"], vel=P["vel"], r_in=P["r_in"], r_out=P["r_out"], ptype=P["ptype"],
                                      theta=theta, n_group_limit=n_group_limit, n_leaf_limit=n_leaf_limit)
    return batch, epi_src, P["prm"], P


# ---------------------------------------------------------------------------------------------
# inputs of the changeover correction
# ---------------------------------------------------------------------------------------------
def corr_particles(P):
    """types.PtclCorr array for the particle set of :func:`kroupa_binary_particles` (same order): members
    carry status < 0 (minus the address of their c.m. particle) and their true mass as backup, the c.m.
    particle status > 0 with the binary mass as backup, probes and orbit samples status > 0 without
    backup, singles status = backup = 0 (reference src/artificial_particles.hpp:20-102)."""
    from .types import PtclCorr
    n, nb = len(P["mass"]), P["n_bin"]
    p = np.zeros(n, dtype=PtclCorr)
    p["id"] = np.arange(1, n + 1)
    p["mass"], p["pos"], p["r_in"], p["r_out"] = P["mass"], P["pos"], P["r_in"], P["r_out"]
    if nb:
        m_star = P.get("m_members")
        cm_adr = 10 * nb + np.arange(nb)
        for half in (0, 1):
            sl = slice(half * nb, (half + 1) * nb)
            p["status"][sl] = -(cm_adr + 1.0)
            p["mass_backup"][sl] = m_star[half] if m_star is not None else 1.0 / P["n_star"]
        p["status"][2 * nb:10 * nb] = 1.0                    # tidal-tensor probes
        p["status"][10 * nb:11 * nb] = 2.0                   # c.m.
        p["mass_backup"][10 * nb:11 * nb] = 3.0 * P["mass"][11 * nb:12 * nb]
        p["status"][11 * nb:14 * nb] = 3.0                   # orbit samples
    return p


def neighbor_lists(pos, rs):
    """CSR neighbour lists with FDPS's symmetric search criterion |x_i - x_j| < max(rs_i, rs_j), the
    particle itself included (as getNeighborListOneParticle returns it), ascending j.  scipy k-d tree."""
    from scipy.spatial import cKDTree
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    rs = np.ascontiguousarray(rs, dtype=np.float64)
    n = len(pos)
    tree = cKDTree(pos)
    hits = tree.query_ball_point(pos, rs, workers=-1, return_sorted=False)   # j within rs_i of i
    cnt = np.fromiter((len(h) for h in hits), dtype=np.int64, count=n)
    ii = np.repeat(np.arange(n, dtype=np.int64), cnt)
    jj = np.fromiter((j for h in hits for j in h), dtype=np.int64, count=int(cnt.sum()))
    d2 = ((pos[ii] - pos[jj]) ** 2).sum(axis=1)
    keep = d2 < rs[ii] ** 2                                                  # strict, as the search kernel tests
    ii, jj = ii[keep], jj[keep]
    key = np.unique(np.concatenate([ii * n + jj, jj * n + ii]))              # symmetrise: j sees i if i sees j
    ii, jj = key // n, key % n
    off = np.zeros(n + 1, dtype=np.int64)
    np.add.at(off, ii + 1, 1)
    off = np.cumsum(off)
    assert off[-1] < 2 ** 31
    return off.astype(np.int32), jj.astype(np.int32)


def plummer_particles(mass, pos, vel, prm):
    """The dict shape of :func:`kroupa_binary_particles` for a set of single stars."""
    r_in, r_out, rs = particle_rout_rsearch(mass, vel, prm)
    return dict(pos=pos, mass=mass, vel=vel, rs=rs, r_in=r_in, r_out=r_out, ptype=np.ones(len(mass), np.int32), prm=prm,
                n_star=len(mass), n_bin=0)


def neighbor_lists_subset(pos, rs, subset):
    """As :func:`neighbor_lists`, for the particles `subset` only: CSR over the subset, j over ALL particles."""
    from scipy.spatial import cKDTree
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    rs = np.ascontiguousarray(rs, dtype=np.float64)
    subset = np.asarray(subset, dtype=np.int64)
    tree = cKDTree(pos)
    hits = tree.query_ball_point(pos[subset], float(rs.max()), workers=-1, return_sorted=True)
    cnt = np.fromiter((len(h) for h in hits), dtype=np.int64, count=len(subset))
    ii = np.repeat(np.arange(len(subset), dtype=np.int64), cnt)
    jj = np.fromiter((j for h in hits for j in h), dtype=np.int64, count=int(cnt.sum()))
    d2 = ((pos[subset[ii]] - pos[jj]) ** 2).sum(axis=1)
    keep = d2 < np.maximum(rs[subset[ii]], rs[jj]) ** 2
    ii, jj = ii[keep], jj[keep]
    off = np.zeros(len(subset) + 1, dtype=np.int64)
    np.add.at(off, ii + 1, 1)
    return np.cumsum(off).astype(np.int32), jj.astype(np.int32)
