"""Times the neighbour-list + changeover-correction extensions at the bench's default
size.  Device: tree_nb-style multiwalk dispatches with option nb_lists, then pb_correct_changeover on those
lists (host buffers in and out); beside it the CPU oracle's correction loop (OpenMP over particles).
    python tools/run_changeover.py [n_star]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from petar_b200 import engine, harness
from petar_b200.types import PtclCorr
from oracle import binding as ob

n_star = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
for kv in sys.argv[2:]:                                      # library options, key=value
    engine.set_option(kv.split("=")[0], int(kv.split("=")[1]))
batch, epi_src, prm, P = harness.kroupa_binary_case(n_star)
out = {"n_particles": int(batch.n_epi_total), "candidate_pairs": float(batch.interactions()[0])}


def med(fn, n=3):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), r


from petar_b200.types import ForceSoft
force = np.zeros(batch.n_epi_total, dtype=ForceSoft)
tables = engine.make_dispatch_tables(batch, force)
engine.tree_neighbor_search(batch, force=force, tables=tables)
out["counts_only_ms"] = 1e3 * med(lambda: engine.tree_neighbor_search(batch, force=force, tables=tables))[0]
engine.tree_neighbor_search(batch, force=force, tables=tables, lists=True)
engine.get_profile(reset=True)
t, (f, off, idx) = med(lambda: engine.tree_neighbor_search(batch, force=force, tables=tables, lists=True))
prof = engine.get_profile()
out["lists_kernels_ms_per_call"] = 1e3 * prof["t_calc"] / 3
out["counts_and_lists_ms"] = 1e3 * t
out["neighbour_pairs"] = int(len(idx))

# correction on the device lists: i-particles in walk order, neighbours = the epj array
allp = harness.corr_particles(P)
pj = np.zeros(len(batch.epj), dtype=PtclCorr)
src = batch.epj["id"] - 1                                   # harness ids are particle index + 1
for k in PtclCorr.names:
    pj[k] = allp[k][src]
pi = allp[epi_src].copy()
for replay in (0, 1):
    ref = pi.copy()
    t0 = time.perf_counter()
    ob.correct_force_tree_neighbor(ref, off, idx, pj, 0.0, prm["r_out"], 1.0, replay)
    t_cpu = time.perf_counter() - t0
    engine.correct_force_with_cutoff_tree_neighbor(pi.copy(), off, idx, pj, 0.0, prm["r_out"], 1.0, replay)
    ts = []
    for _ in range(5):
        q = pi.copy()
        t0 = time.perf_counter()
        engine.correct_force_with_cutoff_tree_neighbor(q, off, idx, pj, 0.0, prm["r_out"], 1.0, replay)
        ts.append(time.perf_counter() - t0)
    out["correction_replay_fp32" if replay else "correction_fp64"] = {
        "device_e2e_ms": 1e3 * float(np.median(ts)), "cpu_oracle_ms": 1e3 * t_cpu, "cpu_threads": os.cpu_count(),
        "bit_identical": bool(q.tobytes() == ref.tobytes())}
print(json.dumps(out))
