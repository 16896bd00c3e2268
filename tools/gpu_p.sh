#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --n-walk-limit 1000 > $O/p_bench_nw1000.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --f-bin 1.0 > $O/p_bench_fbin100.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/p_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/p_ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 60 -c 2 -f -o $O/p_prof_force python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/p_ncu_full_run.log 2>&1
