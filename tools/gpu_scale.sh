#!/bin/bash
# 1 -> 8 GPU scaling run, the way the driver launches it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi -L > $O/s_gpus.log; nproc >> $O/s_gpus.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py 200000 > $O/s_check_2gpu.log 2>&1
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 5 --warmup 3 > $O/s_bench_${n}gpu.log 2>&1
done
python bench.py --gpus 1 --steps 5 --warmup 3 > $O/s_bench_1gpu.log 2>&1
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/s_ref.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 5 --warmup 3 --workload plummer > $O/s_bench_8gpu_plummer.log 2>&1
