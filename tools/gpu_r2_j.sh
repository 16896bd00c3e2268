#!/bin/bash
# round 2, call J: raw_upload A/B at N = 1 (timeline), loop over 6 steps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for ru in 0 1; do
timeout 300 python - $ru <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from petar_b200 import engine, harness as hz
ru=int(sys.argv[1])
batch, _, prm, _ = hz.kroupa_binary_case(400000)
cells, groups = batch.tree.export_tree(out=engine.tree_stage(batch.tree.n_nodes, batch.n_walk))
f = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)
engine.set_option("raw_upload", ru)
for k in range(6):
    t0=time.perf_counter(); engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], force=f, resident=True); dt=time.perf_counter()-t0
    print("raw_upload", ru, "step", k, "wall %.2f ms" % (dt*1e3), {k2: round(v,3) for k2,v in engine.tree_timeline().items()}, flush=True)
PY
done
