#!/bin/bash
# round 2, call N: two i-particles per lane in the SP loops of both force kernels (option sp2i): all GPU tests, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python -m pytest tests -m gpu -q -x > $O/n_pytest.log 2>&1; tail -4 $O/n_pytest.log
for v in 1 0; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --opt sp2i=$v > $O/n_bench_sp2i$v.log 2>&1
python - $O/n_bench_sp2i$v.log <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d=json.loads(line); ef=d["e2e_functors"]
        print(sys.argv[1].split('/')[-1], "value %.1f frac %.3f | tree e2e %.2f ms %s | functors %.2f ms" % (d["value"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["device_timeline_ms_max_over_ranks"], ef["ms_per_step"]))
PY
done
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/n_racecheck.log 2>&1; tail -2 $O/n_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/n_memcheck.log 2>&1; tail -2 $O/n_memcheck.log
