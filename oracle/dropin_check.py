"""Checker for the DROP-IN combination — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py parity).

What an unmodified PeTar integrates is not the kernel's output but `kernel force + CPU changeover correction`
(reference src/petar.hpp:965-1032 -> src/hard.hpp:1655-1691 -> :1408-1476).  With USE_GPU the correction removes the
kernel's clamped neighbour terms by re-computing them in float from absolute float-cast positions, so a kernel is a
correct drop-in iff

    kernel force  +  float-replay correction   ==   fp64 NoSimd force  +  all-double correction

within tolerance.  Both corrections are the oracle's restatement (oracle_changeover.c), which tests/test_oracle.py pins
bit for bit against the reference's own function compiled with and without -DUSE_GPU.
"""
import numpy as np

from petar_b200 import harness as hz
from . import binding as ob


class DropinChecker:
    """Neighbour lists (FDPS's search: |dx| < max of the 0.99 r_search, src/ptcl.hpp:8) and correction inputs for a
    particle set of harness.kroupa_binary_particles / a plain Plummer set; `subset` = particle indices to check."""

    def __init__(self, P, prm, subset=None, p0=None, rs=None):
        """Either P (particle dict) or p0 (PtclCorr rows of the particle set, e.g. what one rank's j store holds: its own
        particles and the LET particles) + rs (their r_search)."""
        self.prm = prm
        self.p0 = hz.corr_particles(P) if p0 is None else np.ascontiguousarray(p0)
        rs = P["rs"] if rs is None else rs
        n = len(self.p0)
        self.subset = np.arange(n) if subset is None else np.asarray(subset)
        if subset is None:
            self.off, self.idx = hz.neighbor_lists(self.p0["pos"], 0.99 * rs)
        else:
            self.off, self.idx = hz.neighbor_lists_subset(self.p0["pos"], 0.99 * rs, self.subset)
        self.n_nb = np.diff(self.off) - 1

    def corrected(self, acc, pot, replay_fp32):
        """acc[len(subset), 3], pot[len(subset)] of the checked particles -> PtclCorr rows after the correction."""
        p = self.p0[self.subset].copy()
        p["acc"], p["pot_tot"], p["pot_soft"] = acc, pot, pot
        prm = self.prm
        return ob.correct_force_tree_neighbor(p, self.off, self.idx, self.p0, prm["eps"], prm["r_out"], prm["G"], replay_fp32)

    def compare(self, f_kernel, f_oracle):
        """f_*: ForceSoft-like arrays (acc, pot) of the checked particles, same order as `subset`."""
        got = self.corrected(f_kernel["acc"], f_kernel["pot"], True)
        want = self.corrected(f_oracle["acc"], f_oracle["pot"], False)
        amag = np.maximum(np.linalg.norm(want["acc"], axis=1), 1e-300)
        ea = np.linalg.norm(got["acc"] - want["acc"], axis=1) / amag
        ep = np.abs(got["pot_tot"] - want["pot_tot"]) / np.maximum(np.abs(want["pot_tot"]), 1e-300)
        has = self.n_nb > 0
        resid = (np.linalg.norm(got["acc"] - want["acc"], axis=1)[has] / self.n_nb[has]) if has.any() else np.zeros(1)
        return {"acc_rel_err": {"median": float(np.median(ea)), "p99": float(np.percentile(ea, 99)), "max": float(ea.max())},
                "pot_tot_rel_err": {"median": float(np.median(ep)), "p99": float(np.percentile(ep, 99)), "max": float(ep.max())},
                "abs_residual_per_neighbour": {"median": float(np.median(resid)), "max": float(resid.max())},
                "n_checked": int(len(ea)), "n_with_neighbours": int(has.sum()), "n_neighbour_pairs": int(self.n_nb.sum()),
                "pass": bool(np.median(ea) <= 1e-6 and ea.max() <= 1e-4 and np.median(ep) <= 1e-6 and ep.max() <= 1e-4)}


plummer_particles = hz.plummer_particles      # (lives with the other input generators)
