#!/bin/bash
# round 2, call O (2 GPUs): LET send rows made on the device (raw lists / raw SP rows), multirank parity, 2-GPU bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_lifecycle.py tests/test_gpu_multirank.py tests/test_gpu_device_walk.py -m gpu -q -x > $O/o_pytest.log 2>&1; tail -4 $O/o_pytest.log
PETAR_B200_TRACE_HOST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 > $O/o_bench_2gpu.log 2>&1
python - $O/o_bench_2gpu.log <<'PY'
import json,sys
ok=False
for fn in sys.argv[1:]:
    for line in open(fn):
        if line.startswith('{"metric"'):
            ok=True
            d=json.loads(line); e=d["e2e"]
            print(fn, "value %.1f (%.2f ms; kernels %.2f xchg %.2f) | e2e %.2f ms | functors %.2f" % (d["value"], d["ms_per_step"], d["value_breakdown"]["kernels_only_ms"], d["value_breakdown"]["let_exchange_ms"], e["ms_per_step"], d["e2e_functors"]["ms_per_step"]))
            print("   timeline", e.get("device_timeline_ms_max_over_ranks")); print("   host", e.get("rank0_host_phases_ms")); print("   let trace", e.get("rank0_let_enqueue_trace_ms"))
            print("   parity", json.dumps(d.get("parity",{}).get("all_ranks")))
    if not ok: print(open(fn).read()[-2500:])
PY
