#!/bin/bash
# A/B of kernel variants built by tools/build_variants.py: parity tests on the default build first,
# then the same bench line per variant.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/x_var_pytest.log 2>&1
tail -3 $O/x_var_pytest.log
for rep in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-device-walk > $O/x_var_default_$rep.log 2>&1
  for d in petar_b200/lib/variants/*/; do
    v=$(basename $d)
    PETAR_B200_LIB=$PWD/$d/libpetar_b200.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-device-walk > $O/x_var_${v}_$rep.log 2>&1
  done
done
for f in $O/x_var_*_[12].log; do echo "$f: $(grep -o '"value": [0-9.e+]*, "unit"' $f | head -1) $(grep -o '"ms_per_step": [0-9.]*' $f)"; done
