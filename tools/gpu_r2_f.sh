#!/bin/bash
# round 2, call F: ncu source-level capture of the persistent force kernel + default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 1 -c 1 -f -o $O/f_prof_persist python tools/run_resident.py 1000000 3 > $O/f_ncu_persist.log 2>&1
ls -la $O/f_prof_persist.ncu-rep; tail -2 $O/f_ncu_persist.log
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > $O/f_bench.log 2>&1; tail -1 $O/f_bench.log | cut -c1-400
