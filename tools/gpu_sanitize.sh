#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
{
echo "compute-sanitizer (CUDA 12.9) over tools/sanitize_small.py — every kernel path (force EP/SP with the TMA index ring, count-only, pair emission + CUB sort, device walk count/fill, reduce, dense field query, changeover correction) on a 3000-star Kroupa+binaries case"
for tool in memcheck racecheck; do
  echo "--- $tool"
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | grep -v "^$" | tail -6
  echo "rc=$?"
done
} > $O/sanitizer.txt 2>&1
cat $O/sanitizer.txt
