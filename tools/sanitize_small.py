"""Small end-to-end invocations of every kernel path, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petar_b200 import engine, harness as hz
from petar_b200.types import EPJSoft

batch, _, prm, _ = hz.kroupa_binary_case(3000)
f = engine.calc_force_all_and_write_back(batch, prm["eps"], prm["r_out"], prm["G"], n_walk_limit=16)
g = engine.tree_neighbor_search(batch, n_walk_limit=16)
cells, groups = batch.tree.export_tree()
h = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])
for ws, sp2i, fuse in ((0, 0, 0), (0, 1, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)):  # device-resident step with every persistent kernel variant: first exact, then speculative
    engine.set_option("ws", ws); engine.set_option("sp2i", sp2i); engine.set_option("fuse_reduce", fuse)
    for _ in range(2):
        hr = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], resident=True)
    assert np.array_equal(hr["n_ngb"], h["n_ngb"])
    assert np.abs(hr["acc"] - h["acc"]).max() <= 1e-5 * np.abs(h["acc"]).max()
part = np.zeros(len(batch.epj), dtype=EPJSoft)
part["pos"], part["mass"] = batch.epj["pos"], batch.epj["mass"]
pts = np.random.default_rng(0).normal(size=(100, 3))
ax, ay, az, phi = engine.get_gravity_and_potential_at_point(pts[:, 0], pts[:, 1], pts[:, 2], part)
assert np.array_equal(f["n_ngb"], h["n_ngb"]) and np.isfinite(ax).all()
# neighbour lists (pair emission + device sort) and the changeover correction on them
from petar_b200.types import PtclCorr
g2, off, idx = engine.tree_neighbor_search(batch, n_walk_limit=16, lists=True)
pj = np.zeros(len(batch.epj), dtype=PtclCorr)
for k in ("id", "mass", "pos", "r_in", "r_out"):
    pj[k] = batch.epj[k]
pi = np.zeros(batch.n_epi_total, dtype=PtclCorr)
pi["id"], pi["pos"], pi["r_in"], pi["r_out"] = batch.epi["id"], batch.epi["pos"], prm["r_in"], prm["r_out"]
for replay in (0, 1):
    c = engine.correct_force_with_cutoff_tree_neighbor(pi.copy(), off, idx, pj, 0.0, prm["r_out"], 1.0, replay)
assert np.array_equal(g2["n_ngb"], g["n_ngb"]) and off[-1] == len(idx) and np.isfinite(c["acc"]).all()
print("sanitize_small ok", f["n_ngb"].sum(), g["n_ngb"].sum(), len(idx))
