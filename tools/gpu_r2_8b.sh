#!/bin/bash
# round 2: 8-GPU host-phase trace of the resident tree step (where do tree staging and LET enqueue spend their time)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
PETAR_B200_TRACE=1 PETAR_B200_TRACE_HOST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 8 --warmup 3 --no-parity > $O/g8b_bench.log 2>&1
grep "petar_b200 trace" $O/g8b_bench.log | head -8
python - $O/g8b_bench.log <<'PY'
import json,sys
for fn in sys.argv[1:]:
    for line in open(fn):
        if line.startswith('{"metric"'):
            d=json.loads(line); e=d["e2e"]
            print(fn, "value %.1f (%.2f ms; kernels %.2f xchg %.2f) | e2e %.2f ms | functors %.2f" % (d["value"], d["ms_per_step"], d["value_breakdown"]["kernels_only_ms"], d["value_breakdown"]["let_exchange_ms"], e["ms_per_step"], d["e2e_functors"]["ms_per_step"]))
            print("   timeline", e.get("device_timeline_ms_max_over_ranks")); print("   host", e.get("rank0_host_phases_ms")); print("   let trace", e.get("rank0_let_enqueue_trace_ms"))
PY
