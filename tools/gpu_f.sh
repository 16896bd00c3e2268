#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams 4 > $O/f_bench_s4.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams 2 > $O/f_bench_s2.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench > $O/f_microbench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/f_launches_1e6.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/f_ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 40 -c 2 -f -o $O/f_prof_force python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/f_ncu_full_run.log 2>&1
