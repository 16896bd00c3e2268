#!/bin/bash
# ncu captures of the extension kernels (device tree walk, neighbour search with pair emission, changeover correction)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -c 2 -f -o $O/x_prof_walk python tools/run_tree_force.py 1000000 1 > $O/x_ncu_walk.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"corr_kernel" -c 1 -f -o $O/x_prof_corr python tools/run_changeover.py 200000 > $O/x_ncu_corr.log 2>&1
ls -la $O/x_prof_walk.ncu-rep $O/x_prof_corr.ncu-rep
