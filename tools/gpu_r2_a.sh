#!/bin/bash
# round 2, call A: loop ceilings + FP32 peak microbench + baseline bench at HEAD
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/a_smi.txt
timeout 120 tools/bin/microbench4 > $O/a_microbench4.txt 2>&1
timeout 300 tools/bin/loopbench > $O/a_loopbench.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/a_bench.log 2>&1
tail -5 $O/a_microbench4.txt; cat $O/a_loopbench.txt
