#!/bin/bash
# round 2: BASELINE config 4 stand-in (N = 1e6 stars, 100 % binaries -> 7e6 tree particles) on one B200, plus the resident tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_device_walk.py -m gpu -q -x > $O/c4_pytest.log 2>&1; tail -3 $O/c4_pytest.log
timeout 1500 python bench.py --f-bin 1.0 --steps 3 --warmup 3 --no-cpu-baseline > $O/c4_bench.log 2>&1
python - $O/c4_bench.log <<'PY'
import json,sys
ok=False
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        ok=True
        d=json.loads(line); e=d["e2e"]; f=d["e2e_functors"]
        print(d["config"]["workload"], d["config"]["n_tree_particles"], "value %.1f (%.2f ms) frac %.3f | tree e2e %.2f ms %s | functors %.2f ms | parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], e["ms_per_step"], e["device_timeline_ms_max_over_ranks"], f["ms_per_step"], json.dumps(d["parity"]["all_ranks"])))
if not ok: print(open(sys.argv[1]).read()[-2000:])
PY
