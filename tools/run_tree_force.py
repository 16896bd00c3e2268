"""Run the device-walk path a few times at N=1e6 (for ncu launch lists / timing breakdowns / option sweeps).
    python tools/run_tree_force.py [n] [reps] [key=value ...]      # options go to pb_set_option"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petar_b200 import engine, harness as hz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batch, _, prm, _ = hz.kroupa_binary_case(n)
cells, groups = batch.tree.export_tree(out=engine.tree_stage(batch.tree.n_nodes, batch.n_walk)) if "--no-stage" not in sys.argv else batch.tree.export_tree()
sys.argv = [a for a in sys.argv if a != "--no-stage"]
sweeps = [a for a in sys.argv[3:]] or [""]
for sw in sweeps:
    for kv in filter(None, sw.split(",")):
        k, v = kv.split("=")
        engine.set_option(k, int(v))
    ts = []
    for k in range(reps):
        engine.get_profile(reset=True)
        t0 = time.perf_counter()
        f = engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"])
        dt = time.perf_counter() - t0
        ts.append(dt)
        p = engine.get_profile()
    print(f"[{sw or 'default'}] median {np.median(ts[1:])*1e3:.1f} ms (last: t_copy {p['t_copy']*1e3:.1f} t_send {p['t_send']*1e3:.1f} t_calc {p['t_calc']*1e3:.1f} "
          f"t_recv {p['t_recv']*1e3:.1f} h2d {p['h2d_bytes']/1e6:.0f} MB launches {p['n_kernel_launch']}) acc checksum {float(np.abs(f['acc']).sum()):.12e}", flush=True)
