#!/bin/bash
# all GPU tests, then three repeats of the kernels/e2e bench line (no CPU legs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nproc
timeout 300 python -m pytest tests -m gpu -x -q > $O/q_pytest.log 2>&1; tail -3 $O/q_pytest.log
B="timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-device-walk"
for rep in 1 2; do $B "$@" > $O/q_bench_$rep.log 2>&1; done
for f in $O/q_bench_[12].log; do python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d=json.loads(line); e=d["e2e"]; r=e["rank0_ms_per_step"]
        print(sys.argv[1].split('/')[-1], "value %.1f e2e %.1f G/s %.2f ms gap %.2f plan %.2f pack %.2f unpack %.2f enq %.2f" % (d["value"], e["value"], e["ms_per_step"], r["gpu_idle_between_walk_groups"], r["host_plan"], r["host_pack"], r["host_unpack"], r["host_enqueue"]))
PY
done
