#!/bin/bash
# e2e A/B on one box, interleaved repeats: sub-batch lead and stream count (the kernels-only value is unaffected)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nproc > $O/y_nproc.log
B="timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-device-walk"
for rep in 1 2 3; do
  $B > $O/y_s4_l1_$rep.log 2>&1
  $B --streams 8 --opt lead=3 > $O/y_s8_l3_$rep.log 2>&1
  $B --streams 8 --opt lead=1 > $O/y_s8_l1_$rep.log 2>&1
  $B --streams 6 --opt lead=2 > $O/y_s6_l2_$rep.log 2>&1
done
for f in $O/y_s*_[123].log; do python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{"metric"'):
        d=json.loads(line); e=d["e2e"]; r=e["rank0_ms_per_step"]
        print(sys.argv[1].split('/')[-1], "value %.1f e2e %.1f G/s %.2f ms gap %.2f plan %.2f pack %.2f unpack %.2f enq %.2f" % (d["value"], e["value"], e["ms_per_step"], r["gpu_idle_between_walk_groups"], r["host_plan"], r["host_pack"], r["host_unpack"], r["host_enqueue"]))
PY
done
