#!/bin/bash
# Round-end verification on one B200: all GPU tests, smoke, the default bench line (both arms), the ncu launch
# list of the same command and one full capture of the force kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 300 python -m pytest tests -m gpu -x -q > $O/f_pytest.log 2>&1; tail -3 $O/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/f_smoke.log 2>&1; tail -1 $O/f_smoke.log
timeout 900 python bench.py > $O/f_bench.log 2>&1; tail -1 $O/f_bench.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/f_bench_ref.log 2>&1; tail -1 $O/f_bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-device-walk > $O/f_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 60 -c 2 -f -o $O/f_prof_force python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-device-walk > $O/f_ncu_full_run.log 2>&1
ls -la $O/f_prof_force.ncu-rep
