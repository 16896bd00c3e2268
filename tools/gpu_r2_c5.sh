#!/bin/bash
# round 2: BASELINE configs[4] — N = 1e7 stars, 100 % binaries (7e7 tree particles), 8 x B200: HBM sizing run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
free -g | head -2 > $O/c5_host.txt; nproc >> $O/c5_host.txt
timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 tools/run_config5.py 10000000 1.0 2 > $O/c5_config5.log 2>&1
grep CONFIG5 $O/c5_config5.log | cut -c1-1500 || tail -30 $O/c5_config5.log
tail -4 $O/c5_config5.log | cut -c1-300; cat $O/c5_host.txt
