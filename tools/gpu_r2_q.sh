#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python tools/run_resident.py 1000000 5 raw_result=1 raw_result=1 > $O/q_resident.log 2>&1; grep "step" $O/q_resident.log
