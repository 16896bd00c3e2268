#!/bin/bash
# round 2, call P: reduction fused into the persistent force kernel; results written into the caller's page-locked array
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 600 python tools/run_resident.py 1000000 5 raw_result=0 raw_result=1 raw_result=0 raw_result=1 > $O/p_resident.log 2>&1; grep "step" $O/p_resident.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/p_pytest.log 2>&1; tail -3 $O/p_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/p_bench.log 2>&1
python - $O/p_bench.log <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d=json.loads(line); ef=d["e2e_functors"]
        print("value %.1f frac %.3f | tree e2e %.2f ms %s | functors %.2f ms | parity %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["device_timeline_ms_max_over_ranks"], ef["ms_per_step"], json.dumps(d["parity"]["all_ranks"])))
PY
