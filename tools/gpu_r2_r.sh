#!/bin/bash
# round 2, call R: fused reduction in the functor path (force_kernel), all GPU tests, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python -m pytest tests -m gpu -q -x > $O/r_pytest.log 2>&1; tail -4 $O/r_pytest.log
for v in 1 0; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --opt fuse_reduce=$v > $O/r_bench_fuse$v.log 2>&1
python - $O/r_bench_fuse$v.log <<'PY'
import json,sys
ok=False
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        ok=True
        d=json.loads(line); ef=d["e2e_functors"]
        print(sys.argv[1].split('/')[-1], "value %.1f frac %.3f | tree e2e %.2f ms | functors %.2f ms" % (d["value"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], ef["ms_per_step"]), {k:round(v,2) for k,v in ef["rank0_ms_per_step"].items() if k!="note"}, d["gpu_launches"])
if not ok: print(open(sys.argv[1]).read()[-1500:])
PY
done
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/r_racecheck.log 2>&1; tail -2 $O/r_racecheck.log
