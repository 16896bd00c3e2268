#!/bin/bash
# round 2: 8-GPU bench lines (per-rank parity, device timeline of the resident tree step), raw_upload on / off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l > $O/g8_gpus.txt; nproc >> $O/g8_gpus.txt
PETAR_B200_RAW_UPLOAD=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 8 --warmup 3 > $O/g8_bench.log 2>&1
PETAR_B200_RAW_UPLOAD=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 8 --steps 8 --warmup 3 --no-parity > $O/g8_bench_raw0.log 2>&1
python - $O/g8_bench.log $O/g8_bench_raw0.log <<'PY'
import json,sys
for fn in sys.argv[1:]:
    ok=False
    for line in open(fn):
        if line.startswith('{"metric"'):
            ok=True
            d=json.loads(line); e=d["e2e"]
            print(fn, "value %.1f (%.2f ms; kernels %.2f xchg %.2f) frac %.3f | e2e %.1f G/s %.2f ms" % (d["value"], d["ms_per_step"], d["value_breakdown"]["kernels_only_ms"], d["value_breakdown"]["let_exchange_ms"], d["roofline"]["frac"], e["value"], e["ms_per_step"]))
            print("    e2e timeline", e.get("device_timeline_ms_max_over_ranks"), e.get("rank0_host_phases_ms"))
            print("    functors %.2f ms" % d["e2e_functors"]["ms_per_step"])
            print("    parity", json.dumps(d.get("parity",{}).get("all_ranks")))
    if not ok: print(open(fn).read()[-2000:])
PY
