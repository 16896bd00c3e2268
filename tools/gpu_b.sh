#!/bin/bash
# second GPU call of round 1: full test log, pipe microbenchmarks, ncu launch list + full capture, option sweeps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m petar_b200.build > $O/b_build.log 2>&1
python -m pytest tests -m gpu -q -s > $O/b_tests.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench > $O/b_microbench.log 2>&1
export PATH=$PATH:/usr/local/cuda/bin
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/b_launches_1e5.csv python bench.py --n 100000 --steps 2 --warmup 1 --no-cpu-baseline > $O/b_ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 6 -c 3 -f -o $O/b_prof_force python bench.py --n 100000 --steps 2 --warmup 1 --no-cpu-baseline > $O/b_ncu_full_run.log 2>&1
for s in 1 2 4 8; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $s > $O/b_sweep_streams$s.log 2>&1; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --nr 1 > $O/b_sweep_nr1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --cull 0 > $O/b_sweep_cull0.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --jchunk 512 > $O/b_sweep_jchunk512.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --jchunk 2048 > $O/b_sweep_jchunk2048.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/b_ref_1e6.log 2>&1
