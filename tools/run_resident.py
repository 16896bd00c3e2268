"""A few device-resident tree steps of the BASELINE config-3 stand-in (N = 1e6) for profilers: python tools/run_resident.py [n] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petar_b200 import engine, harness as hz
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
batch, _, prm, _ = hz.kroupa_binary_case(n)
cells, groups = batch.tree.export_tree(out=engine.tree_stage(batch.tree.n_nodes, batch.n_walk))
f = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)
for _ in range(steps):
    engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], force=f, resident=True)
print("timeline", engine.tree_timeline())
