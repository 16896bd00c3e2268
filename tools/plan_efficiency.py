"""Lane x time efficiency of the task plan on the config-3 stand-in (CPU only; the model quoted in DESIGN §4.1/§4.3).

A walk with n_i i-particles occupies ceil(n_i / 32) i-blocks of 32 lanes for the whole of its EP and SP lists; the ragged last
block lets 2 (<= 16 particles left) or 4 (<= 8) lanes share a particle and split the j pairs.  j side: pairs of j, the last tile of
a list is split evenly between the warps that share the i-block.  Weights: 1 per EP entry, 65/38 per SP entry (the kernel's
measured cost ratio is ~2)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from petar_b200 import harness as hz

def main(n=1000000):
    batch, epi_src, prm, P = hz.kroupa_binary_case(n, 0.1)
    ni = np.diff(batch.i_off); nej = np.diff(batch.ej_off); nsj = np.diff(batch.sj_off)
    nb = (ni + 31) // 32
    rem = ni - 32 * (nb - 1)
    share = np.where(rem <= 8, 4, np.where(rem <= 16, 2, 1))
    used = 32 * (nb - 1) + rem * share
    work = nej + 2.0 * nsj
    e_i = float((work * used).sum() / (work * nb * 32).sum())
    e_i_noshare = float((work * ni).sum() / (work * nb * 32).sum())
    # j side: a lane-sharing group of s lanes takes ceil(pairs / s) steps per tile; pairs = ceil(n / 2)
    def jeff(nj):
        pairs = (nj + 1) // 2
        return float(nj.sum() / (2.0 * pairs).sum())
    out = {"n_walks": int(len(ni)), "mean_i_per_walk": float(ni.mean()), "mean_ep_per_walk": float(nej.mean()), "mean_sp_per_walk": float(nsj.mean()),
           "i_lane_efficiency": e_i, "i_lane_efficiency_without_lane_sharing": e_i_noshare,
           "j_pair_efficiency": {"ep": jeff(nej), "sp": jeff(nsj)},
           "histogram_i_per_walk": {str(k): int(((ni > k - 32) & (ni <= k)).sum()) for k in range(32, 513, 32)}}
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main(int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000)
