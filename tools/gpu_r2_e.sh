#!/bin/bash
# round 2, call E (2 GPUs): multi-rank parity test on real NCCL, bench at N = 2 and N = 1 with the restructured bench.py
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
nvidia-smi -L > $O/e_gpus.txt; nproc >> $O/e_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s > $O/e_pytest_multirank.log 2>&1; tail -12 $O/e_pytest_multirank.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/e_bench_2gpu.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > $O/e_bench_1gpu.log 2>&1
for f in 2gpu 1gpu; do python - $O/e_bench_$f.log <<'PY'
import json,sys
ok=False
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        ok=True
        d=json.loads(line); e=d["e2e"]
        print(sys.argv[1].split('/')[-1], "value %.1f (%.2f ms; kernels %.2f xchg %.2f) frac %.3f | e2e %.1f G/s %.2f ms" % (d["value"], d["ms_per_step"], d["value_breakdown"]["kernels_only_ms"], d["value_breakdown"]["let_exchange_ms"], d["roofline"]["frac"], e["value"], e["ms_per_step"]))
        for k in ("e2e_functors","device_walk"):
            if k in d: print("   ",k, "%.2f ms" % d[k]["ms_per_step"], d[k].get("device_timeline_ms_max_over_ranks"), d[k].get("rank0_host_phases_ms"))
        print("    e2e timeline", e.get("device_timeline_ms_max_over_ranks"), e.get("rank0_host_phases_ms"))
        print("    parity", json.dumps(d.get("parity"))[:900])
if not ok: print(open(sys.argv[1]).read()[-3000:])
PY
done
