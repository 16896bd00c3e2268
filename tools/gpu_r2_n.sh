#!/bin/bash
# round 2: bench line at N GPUs (N = number of GPUs of the box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 8 --warmup 3 > $O/s_bench_${N}gpu.log 2>&1
python - $O/s_bench_${N}gpu.log <<'PY'
import json,sys
for fn in sys.argv[1:]:
    ok=False
    for line in open(fn):
        if line.startswith('{"metric"'):
            ok=True
            d=json.loads(line); e=d["e2e"]
            print(fn, "value %.1f (%.2f ms; kernels %.2f xchg %.2f) frac %.3f | e2e %.1f G/s %.2f ms | functors %.2f" % (d["value"], d["ms_per_step"], d["value_breakdown"]["kernels_only_ms"], d["value_breakdown"]["let_exchange_ms"], d["roofline"]["frac"], e["value"], e["ms_per_step"], d["e2e_functors"]["ms_per_step"]))
            print("    e2e timeline", e.get("device_timeline_ms_max_over_ranks"), e.get("rank0_host_phases_ms"))
            print("    parity", json.dumps(d.get("parity",{}).get("all_ranks")))
    if not ok: print(open(fn).read()[-2000:])
PY
