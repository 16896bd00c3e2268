#!/bin/bash
# round 2, call G: task size sweep of the persistent (device-planned) force launch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for j in 1024 1536 2048 3072 4096 6144 8192; do
python - $j <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from petar_b200 import engine, harness as hz
j=int(sys.argv[1])
if not hasattr(sys.modules[__name__], "_c"):
    pass
batch, _, prm, _ = hz.kroupa_binary_case(1000000)
cells, groups = batch.tree.export_tree(out=engine.tree_stage(batch.tree.n_nodes, batch.n_walk))
f = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)
engine.set_option("jchunk", j)
for _ in range(4):
    engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], force=f, resident=True)
print("jchunk", j, engine.tree_timeline(), flush=True)
PY
done
