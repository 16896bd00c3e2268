#!/bin/bash
# round 2, call D: resident (device-planned, persistent) tree step + mode-2 refinements
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python -m pytest tests/test_gpu_device_walk.py tests/test_gpu_replay_gap.py tests/test_gpu_fullsize.py -m gpu -q -s > $O/d_pytest.log 2>&1; tail -25 $O/d_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/d_bench_default.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --opt coords=0 > $O/d_bench_coords0.log 2>&1
for f in default coords0; do python - $O/d_bench_$f.log <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d=json.loads(line); e=d["e2e"]
        print(sys.argv[1].split('/')[-1], "value %.1f (%.2f ms) frac %.3f e2e %.1f G/s %.2f ms" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], e["value"], e["ms_per_step"]))
        print("   device_walk", d.get("device_walk"))
PY
done
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/d_memcheck.log 2>&1; tail -3 $O/d_memcheck.log
