#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi -L > $O/e_gpus.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams 4 --occupancy 3 > $O/e_bench_occ3.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams 4 --occupancy 2 > $O/e_bench_occ2.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 5 --warmup 3 --streams 4 > $O/e_bench_2gpu.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 tools/check_multigpu.py > $O/e_check_2gpu.log 2>&1
