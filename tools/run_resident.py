"""A few device-resident tree steps of the BASELINE config-3 stand-in (N = 1e6) for profilers and A/B runs:
python tools/run_resident.py [n] [steps] [key=value ...]   (options of pb_set_option, applied in turn: one timeline line per setting)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petar_b200 import engine, harness as hz
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
opts = [a.split("=") for a in sys.argv[3:]] or [None]
batch, _, prm, _ = hz.kroupa_binary_case(n)
cells, groups = batch.tree.export_tree(out=engine.tree_stage(batch.tree.n_nodes, batch.n_walk))
f = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)
ref = None
for o in opts:
    if o: engine.set_option(o[0], int(o[1]))
    tl = {k: 0.0 for k in engine.TIMELINE_KEYS}; wall = 0.0
    for it in range(steps + 2):
        t0 = time.perf_counter()
        engine.tree_force(batch, cells, groups, prm["eps"], prm["r_out"], prm["G"], force=f, resident=True)
        if it >= 2:
            wall += time.perf_counter() - t0
            for k, v in engine.tree_timeline().items(): tl[k] += v
    msg = ""
    if ref is None: ref = f.copy()
    else:
        a0 = np.linalg.norm(ref["acc"], axis=1)
        msg = " | max rel acc diff vs first setting %.2e, n_ngb equal %s" % (float((np.linalg.norm(f["acc"] - ref["acc"], axis=1) / a0).max()), bool((f["n_ngb"] == ref["n_ngb"]).all()))
    print("%s: step %.2f ms, timeline" % ("=".join(o) if o else "default", 1e3 * wall / steps), {k: round(v / steps, 3) for k, v in tl.items()}, msg, flush=True)
