import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petar_b200 import engine, harness as hz
from petar_b200.walks import WalkBatch
from oracle import binding as ob

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
batch, _, prm, _ = hz.plummer_case(n)
eps, r_out, G = prm["eps"], prm["r_out"], prm["G"]
ref = ob.walks_index(batch, eps, r_out, G)
f = engine.calc_force_all_and_write_back(batch, eps, r_out, G)
ea = np.linalg.norm(f["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
print("median", np.median(ea), "max", ea.max(), "n>1e-4:", (ea > 1e-4).sum(), "n>1e-5:", (ea > 1e-5).sum())
worst = np.argsort(-ea)[:12]
# EP-only / SP-only references
none_s = WalkBatch(batch.epj, batch.spj, batch.epi, batch.i_off, batch.id_epj, batch.ej_off, batch.id_spj[:0], np.zeros(batch.n_walk + 1, dtype=np.int64))
f_ep = engine.calc_force_all_and_write_back(none_s, eps, r_out, G)
r_ep = ob.walks_index(none_s, eps, r_out, G)
ea_ep = np.linalg.norm(f_ep["acc"] - r_ep["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
print("EP-only err (normalised by total |a|): median", np.median(ea_ep), "max", ea_ep.max())
for k in worst:
    w = np.searchsorted(batch.i_off, k, side="right") - 1
    il = k - batch.i_off[w]
    ni, ne, ns = batch.n_epi[w], batch.n_epj[w], batch.n_spj[w]
    e = batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]]
    d = batch.epj["pos"][e] - batch.epi["pos"][k]
    r = np.sqrt((d * d).sum(1))
    rs = np.sort(r)[:4]
    print(f"i={k} walk={w} il={il} ni={ni} nej={ne} nsj={ns} err={ea[k]:.3e} err_ep={ea_ep[k]:.3e} nngb gpu/ref {f['n_ngb'][k]}/{ref['n_ngb'][k]} |a|={np.linalg.norm(ref['acc'][k]):.3e} nearest r={rs} r_out={r_out:.3e}")
    print("    gpu", f["acc"][k], "ref", ref["acc"][k], "pos", batch.epi["pos"][k])
# error vs lane / group position
il_all = np.arange(batch.n_epi_total) - np.repeat(batch.i_off[:-1], batch.n_epi)
bad = ea > 1e-4
print("bad count by il//32:", np.bincount((il_all[bad] // 32).astype(int)))
print("bad walks ni:", np.unique(np.repeat(batch.n_epi, batch.n_epi)[bad])[:40])
# nearest-neighbour distance of bad ones vs all
