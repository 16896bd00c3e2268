#!/bin/bash
# round 2, call C: GPU tests with the mbarrier tile pipeline, replay-gap report, A/B bench lines, sanitizer on a small case
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python -m pytest tests -m gpu -q > $O/c_pytest.log 2>&1; tail -15 $O/c_pytest.log
timeout 300 python tests/test_gpu_replay_gap.py > $O/c_replay_gap.json 2>&1; tail -70 $O/c_replay_gap.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c_smoke.log 2>&1; tail -2 $O/c_smoke.log
B="timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-device-walk"
$B > $O/c_bench_default.log 2>&1
$B --opt coords=0 > $O/c_bench_coords0.log 2>&1
$B --opt chunk_tile=0 > $O/c_bench_chunk8.log 2>&1
for f in default coords0 chunk8; do python - $O/c_bench_$f.log <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d=json.loads(line); e=d["e2e"]
        print(sys.argv[1].split('/')[-1], "value %.1f (%.2f ms) frac %.3f e2e %.1f G/s %.2f ms" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], e["value"], e["ms_per_step"]))
PY
done
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/c_racecheck.log 2>&1; tail -3 $O/c_racecheck.log
