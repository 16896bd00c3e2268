#!/bin/bash
# 2-GPU check at HEAD: per-rank parity through NCCL LET exchange, then the 2-GPU and 1-GPU bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py 200000 > $O/t_check_2gpu.log 2>&1; tail -2 $O/t_check_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 > $O/t_bench_2gpu.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-device-walk > $O/t_bench_1gpu.log 2>&1
python tools/summ.py $O/t_bench_2gpu.log $O/t_bench_1gpu.log | cut -c1-330
grep -o '"clocks": {[^}]*}' $O/t_bench_2gpu.log $O/t_bench_1gpu.log
