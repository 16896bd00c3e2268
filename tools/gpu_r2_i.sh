#!/bin/bash
# round 2, call I (2 GPUs): compact walk + raw upload: tests, timeline A/B at N = 1, bench at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python -m pytest tests/test_gpu_device_walk.py tests/test_gpu_lifecycle.py tests/test_gpu_multirank.py tests/test_gpu_parity.py -m gpu -q -x > $O/i_pytest.log 2>&1; tail -6 $O/i_pytest.log
for wc in 1 0; do
timeout 300 python - $wc <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from petar_b200 import engine, harness as hz
wc=int(sys.argv[1])
batch, _, prm, _ = hz.kroupa_binary_case(1000000)
cells, groups = batch.tree.export_tree(out=engine.tree_stage(batch.tree.n_nodes, batch.n_walk))
f = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)
engine.set_option("walk_compact", wc)
import time
for _ in range(4):
    t0=time.perf_counter(); engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], force=f, resident=True); dt=time.perf_counter()-t0
print("walk_compact", wc, "wall %.2f ms" % (dt*1e3), engine.tree_timeline(), flush=True)
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/i_bench_2gpu.log 2>&1
python - $O/i_bench_2gpu.log <<'PY'
import json,sys
ok=False
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        ok=True
        d=json.loads(line); e=d["e2e"]
        print("value %.1f (%.2f ms; kernels %.2f xchg %.2f) | e2e %.1f G/s %.2f ms" % (d["value"], d["ms_per_step"], d["value_breakdown"]["kernels_only_ms"], d["value_breakdown"]["let_exchange_ms"], e["value"], e["ms_per_step"]))
        print("    e2e timeline", e.get("device_timeline_ms_max_over_ranks"), e.get("rank0_host_phases_ms"))
        print("    functors %.2f ms" % d["e2e_functors"]["ms_per_step"]); print("    parity", json.dumps(d.get("parity",{}).get("all_ranks")))
if not ok: print(open(sys.argv[1]).read()[-3000:])
PY
