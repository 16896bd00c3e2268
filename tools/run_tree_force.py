"""Run the device-walk path once or twice at N=1e6 (for ncu launch lists / timing breakdowns)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petar_b200 import engine, harness as hz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batch, _, prm, _ = hz.kroupa_binary_case(n)
cells, groups = batch.tree.export_tree()
for k in range(reps):
    engine.get_profile(reset=True)
    t0 = time.perf_counter()
    f = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])
    dt = time.perf_counter() - t0
    p = engine.get_profile()
    print(f"rep {k}: {dt*1e3:.1f} ms  t_copy {p['t_copy']*1e3:.1f} t_send {p['t_send']*1e3:.1f} t_calc {p['t_calc']*1e3:.1f} t_recv {p['t_recv']*1e3:.1f} "
          f"h2d {p['h2d_bytes']/1e6:.0f} MB launches {p['n_kernel_launch']}", flush=True)
