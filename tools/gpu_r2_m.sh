#!/bin/bash
# round 2, call M: WS kernel variants: ws = 1 (round-2 base), 2 (two i per lane in SP), 3 (2 + per-warp partial slots, no task-end barrier)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 600 python tools/run_resident.py 1000000 5 ws=1 ws=2 ws=3 ws=2 ws=3 > $O/m_resident.log 2>&1; grep "step" $O/m_resident.log
timeout 600 python tools/run_resident.py 20000 2 ws=0 ws=3 > $O/m_resident_small.log 2>&1; grep "step" $O/m_resident_small.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/m_racecheck.log 2>&1; tail -2 $O/m_racecheck.log
