#!/bin/bash
# 8- and 4-GPU bench lines, the way the driver launches them
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l; nproc
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 5 --warmup 3 > $O/u_bench_${n}gpu.log 2>&1
done
python tools/summ.py $O/u_bench_8gpu.log $O/u_bench_4gpu.log | cut -c1-400
