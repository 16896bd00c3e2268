#!/bin/bash
# round 2, call H: warp-specialised persistent kernel: tests, racecheck, A/B timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 600 python -m pytest tests/test_gpu_device_walk.py -m gpu -q -x > $O/h_pytest.log 2>&1; tail -5 $O/h_pytest.log
for ws in 1 0; do
timeout 300 python - $ws <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from petar_b200 import engine, harness as hz
ws=int(sys.argv[1])
batch, _, prm, _ = hz.kroupa_binary_case(1000000)
cells, groups = batch.tree.export_tree(out=engine.tree_stage(batch.tree.n_nodes, batch.n_walk))
f = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)
engine.set_option("ws", ws)
for _ in range(4):
    engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], force=f, resident=True)
print("ws", ws, engine.tree_timeline(), flush=True)
PY
done
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/h_racecheck.log 2>&1; tail -3 $O/h_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/h_memcheck.log 2>&1; tail -3 $O/h_memcheck.log
