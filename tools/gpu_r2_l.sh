#!/bin/bash
# round 2, call L: functor path with pipelined run scan + faster expansion: lifecycle tests, bench A/B ep_runs, ncu capture of force_kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 600 python -m pytest tests/test_gpu_lifecycle.py tests/test_gpu_parity.py tests/test_gpu_neighbor_search.py -m gpu -q -x > $O/l_pytest.log 2>&1; tail -4 $O/l_pytest.log
for er in 1 0; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --opt ep_runs=$er > $O/l_bench_runs$er.log 2>&1
python - $O/l_bench_runs$er.log <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d=json.loads(line); ef=d["e2e_functors"]
        print(sys.argv[1].split('/')[-1], "value %.1f | tree e2e %.2f ms | functors %.2f ms h2d %.0f MB" % (d["value"], d["e2e"]["ms_per_step"], ef["ms_per_step"], ef["h2d_bytes_per_step"]/1e6), {k:round(v,2) for k,v in ef["rank0_ms_per_step"].items() if k!="note"})
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^force_kernel$' -s 60 -c 3 -f -o $O/z_prof_force python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-device-walk > $O/z_ncu_full_run.log 2>&1
ls -la $O/z_prof_force.ncu-rep; tail -3 $O/z_ncu_full_run.log | cut -c1-200
