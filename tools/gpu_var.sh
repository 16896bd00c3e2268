#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-device-walk > $O/x_var_default.log 2>&1
for v in ep2_sp2 ep8_sp2 ep4_sp1 ep4_sp4 ep8_sp4 ep2_sp4; do
  PETAR_B200_LIB=$PWD/petar_b200/lib/variants/$v/libpetar_b200.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-device-walk > $O/x_var_$v.log 2>&1
done
