#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 5 --warmup 3 > $O/j_bench_4gpu.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/j_bench_1gpu_lead1.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --opt lead=0 > $O/j_bench_1gpu_lead0.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams 6 > $O/j_bench_1gpu_s6.log 2>&1
