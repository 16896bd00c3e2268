#!/bin/bash
# Round-2 verification on one B200: all GPU tests, smoke, the default bench line (both arms), the ncu launch list of the
# same command (whole steps: -c large enough) and full captures of the two force kernels and the walk; sanitizer pass.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python -m pytest tests -m gpu -q > $O/z_pytest.log 2>&1; tail -4 $O/z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/z_smoke.log 2>&1; tail -2 $O/z_smoke.log | cut -c1-400
timeout 900 python bench.py > $O/z_bench.log 2>&1; tail -1 $O/z_bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/z_bench_ref.log 2>&1; tail -1 $O/z_bench_ref.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --opt coords=0 > $O/z_bench_coords0.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/z_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/z_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^force_kernel$' -s 60 -c 3 -f -o $O/z_prof_force python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-device-walk > $O/z_ncu_full_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_kernel_ws -s 1 -c 1 -f -o $O/z_prof_ws python tools/run_resident.py 1000000 1 > $O/z_ncu_ws_run.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"walk_kernel_c" -s 1 -c 1 -f -o $O/z_prof_walk python tools/run_resident.py 1000000 1 > $O/z_ncu_walk_run.log 2>&1
ls -la $O/z_prof_force.ncu-rep $O/z_prof_ws.ncu-rep $O/z_prof_walk.ncu-rep
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/z_memcheck.log 2>&1; tail -2 $O/z_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/z_racecheck.log 2>&1; tail -2 $O/z_racecheck.log
