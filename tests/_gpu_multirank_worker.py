"""Worker of tests/test_gpu_multirank.py (run under torchrun, one rank per GPU): parity of the multi-GPU paths on real
NCCL — every rank's forces against the fp64 oracle over the same lists, for the functor path (host lists) and the
device-resident tree step (lists, plan and forces on the GPU; LET EP rows gathered on the device), in the library's
default mode (coords = 2, checked on the corrected force) and at kernel level (coords = 0)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_w = int(os.environ.get("WORLD_SIZE", "1"))
if os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _w))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from petar_b200 import engine, harness as hz, multigpu  # noqa: E402
from oracle import binding as ob  # noqa: E402
from oracle.dropin_check import DropinChecker  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    L = engine.load()
    engine.check(L.pb_init(rank, lr), "pb_init")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    P = hz.kroupa_binary_particles(n, f_bin=0.1, seed=1)
    prm = P["prm"]
    wl = multigpu.build_domain_workload(P["pos"], P["mass"], P["vel"], P["rs"], P["r_in"], P["r_out"], rank, world, dist, ptype=P["ptype"])
    wl["prm"] = prm
    st = multigpu.DomainStepper(wl, rank, world, dist)
    b = wl["batch"]
    ref = ob.walks_index(b, prm["eps"], prm["r_out"], prm["G"])
    gid = wl["store_gid"]
    chk = DropinChecker(None, prm, subset=wl["epi_src"], p0=hz.corr_particles(P)[gid], rs=P["rs"][gid])
    out, ok = {"rank": rank, "world": world, "n_loc": int(wl["n_loc"]), "let_ep": int(wl["n_let_ep"]), "let_sp": int(wl["n_let_sp"]),
               "nccl_bytes": int(st.nccl_bytes_per_step)}, True

    def kernel_stats(f):
        ea = np.linalg.norm(f["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
        ep = np.abs((f["pot"] - ref["pot"]) / ref["pot"])
        good = bool(np.median(ea) <= 1e-6 and ea.max() <= 1e-4 and np.median(ep) <= 1e-6 and ep.max() <= 1e-4 and np.array_equal(f["n_ngb"], ref["n_ngb"]))
        return {"acc_median": float(np.median(ea)), "acc_max": float(ea.max()), "pot_max": float(ep.max()), "pass": good}

    for name, step in (("functors", st.step), ("resident_tree_step", st.step_device_walk)):
        f = np.zeros(b.n_epi_total, dtype=engine.ForceSoft)
        assert engine.get_option("coords") == 2
        for _ in range(3):                                   # the third resident step is speculative (no host round trip)
            step(f)
        rep = chk.compare(f, ref)
        rep["n_ngb_equal"] = bool(np.array_equal(f["n_ngb"], ref["n_ngb"]))
        engine.set_option("coords", 0)
        try:
            f0 = np.zeros_like(f)
            for _ in range(2):
                step(f0)
        finally:
            engine.set_option("coords", 2)
        out[name] = {"dropin_coords2_plus_float_replay": {k: rep[k] for k in ("acc_rel_err", "pot_tot_rel_err", "pass", "n_ngb_equal")},
                     "kernel_coords0": kernel_stats(f0)}
        ok = ok and rep["pass"] and rep["n_ngb_equal"] and out[name]["kernel_coords0"]["pass"]
    out["timeline_ms"] = engine.tree_timeline()
    out["pass"] = bool(ok)
    print("MULTIRANK " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
