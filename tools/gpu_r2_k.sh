#!/bin/bash
# round 2, call K (2 GPUs): config-5 script at a small size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/run_config5.py 200000 1.0 2 > $O/k_config5_small.log 2>&1
grep CONFIG5 $O/k_config5_small.log | cut -c1-1800 || tail -30 $O/k_config5_small.log
tail -5 $O/k_config5_small.log | cut -c1-300
