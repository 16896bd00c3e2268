#!/bin/bash
# round 2, call B: all GPU tests at HEAD (coords = 2 default, tile-aligned chunks), replay-gap report, bench, ncu source capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 600 python -m pytest tests -m gpu -x -q > $O/b_pytest.log 2>&1; tail -5 $O/b_pytest.log
timeout 300 python tests/test_gpu_replay_gap.py > $O/b_replay_gap.json 2>&1; tail -60 $O/b_replay_gap.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/b_bench.log 2>&1; tail -1 $O/b_bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 60 -c 3 -f -o $O/b_prof_force python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-device-walk > $O/b_ncu_full_run.log 2>&1
ls -la $O/b_prof_force.ncu-rep
