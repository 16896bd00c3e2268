"""Times the changeover correction at the bench's default size: device
(pb_correct_changeover, host buffers in and out) beside the CPU oracle with OpenMP over particles.
    python tools/run_changeover.py [n_star]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from petar_b200 import engine, harness
from oracle import binding as ob

n_star = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
P = harness.kroupa_binary_particles(n_star)
p = harness.corr_particles(P)
t0 = time.perf_counter()
off, idx = harness.neighbor_lists(P["pos"], P["rs"])
t_lists = time.perf_counter() - t0
r_out = P["prm"]["r_out"]
out = {"n_particles": len(p), "neighbour_pairs": int(len(idx) - len(p)), "lists_host_s": t_lists}
for replay in (0, 1):
    ref = p.copy()
    t0 = time.perf_counter()
    ob.correct_force_tree_neighbor(ref, off, idx, p, 0.0, r_out, 1.0, replay)
    t_cpu = time.perf_counter() - t0
    engine.correct_force_with_cutoff_tree_neighbor(p.copy(), off, idx, p, 0.0, r_out, 1.0, replay)      # warm-up (allocations)
    ts = []
    for _ in range(5):
        q = p.copy()
        t0 = time.perf_counter()
        engine.correct_force_with_cutoff_tree_neighbor(q, off, idx, p, 0.0, r_out, 1.0, replay)
        ts.append(time.perf_counter() - t0)
    out["replay_fp32" if replay else "fp64"] = {"device_e2e_ms": 1e3 * float(np.median(ts)), "cpu_oracle_ms": 1e3 * t_cpu,
                                                 "cpu_threads": os.cpu_count(), "bit_identical": bool(q.tobytes() == ref.tobytes())}
print(json.dumps(out))
