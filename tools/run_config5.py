"""BASELINE configs[4]: N = 1e7 stars, 100 % binaries, 8 x B200 — HBM sizing run of the device-resident tree step.

    torchrun --nproc-per-node 8 tools/run_config5.py [n_star=10000000] [f_bin=1.0] [steps=2]

Every rank builds its own domain (harness: trees and i-groups only — host interaction lists of 7e7 tree particles would
not fit, and the device builds them anyway), uploads tree + own particles, exchanges the LET over NCCL and runs
pb_tree_force_resident.  Rank 0 prints one JSON line: per-rank particle / cell / list counts, device memory used, step
time and device timeline, and a spot check of a few i-groups against the fp64 oracle (their lists walked on the host).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_w = int(os.environ.get("WORLD_SIZE", "1"))
if os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _w))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from petar_b200 import engine, harness as hz, multigpu  # noqa: E402
from petar_b200.walks import WalkBatch  # noqa: E402


def main():
    n_star = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    f_bin = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    L = engine.load()
    engine.check(L.pb_init(rank, lr), "pb_init")
    t0 = time.time()
    P = hz.kroupa_binary_particles(n_star, f_bin=f_bin, seed=1)
    prm = P["prm"]
    t_gen = time.time() - t0
    hz.skip_walks(True)
    wl = multigpu.build_domain_workload(P["pos"], P["mass"], P["vel"], P["rs"], P["r_in"], P["r_out"], rank, world, dist, ptype=P["ptype"])
    hz.skip_walks(False)
    wl["prm"] = prm
    n_tree = len(P["mass"])
    ptype_all, rs_all = P["ptype"], P["rs"]
    del P
    t_build = time.time() - t0
    b = wl["batch"]
    free0, total = torch.cuda.mem_get_info()
    st = multigpu.DomainStepper(wl, rank, world, dist)
    f = np.zeros(b.n_epi_total, dtype=engine.ForceSoft)
    st.step_device_walk(f)                                   # exact first step (sizes everything)
    torch.cuda.synchronize(); dist.barrier()
    engine.get_profile(reset=True)
    t1 = time.perf_counter()
    for _ in range(steps):
        st.step_device_walk(f)                               # speculative steps: no host round trip
    torch.cuda.synchronize(); dist.barrier()
    sec = (time.perf_counter() - t1) / steps
    prof = engine.get_profile()
    tl = engine.tree_timeline()
    free1, _ = torch.cuda.mem_get_info()

    # spot check: a few groups against the fp64 oracle, lists walked on the host now
    from oracle import binding as ob
    t = b.tree
    epj_src = t.export()[0]
    sp_src = t.let_sp_src()
    n_nodes = wl["n_nodes"]
    rng = np.random.default_rng(rank)
    gs = np.sort(rng.choice(b.n_walk, min(6, b.n_walk), replace=False))
    ea, ep, nbad = [], [], 0
    engine.set_option("coords", 0)
    st.step_device_walk(f)                                   # kernel-level comparison: walk-relative dx for every pair
    engine.set_option("coords", 2)
    for g in gs:
        e, s = t.walk_group(int(g))
        e = epj_src[e].astype(np.int32)
        s = np.where(s < n_nodes, s, n_nodes + sp_src[np.maximum(s - n_nodes, 0)]).astype(np.int32) if len(sp_src) else s
        i0, i1 = int(b.i_off[g]), int(b.i_off[g + 1])
        sub = WalkBatch(b.epj, b.spj, b.epi[i0:i1], [0, i1 - i0], e, [0, len(e)], s, [0, len(s)])
        ref = ob.walks_index(sub, prm["eps"], prm["r_out"], prm["G"])
        got = f[i0:i1]
        ea.append(np.linalg.norm(got["acc"] - ref["acc"], axis=1) / np.maximum(np.linalg.norm(ref["acc"], axis=1), 1e-300))
        ep.append(np.abs(got["pot"] - ref["pot"]) / np.maximum(np.abs(ref["pot"]), 1e-300))
        nbad += int((got["n_ngb"] != ref["n_ngb"]).sum())
    ea, ep = np.concatenate(ea), np.concatenate(ep)
    mine = {"rank": rank, "n_loc": int(wl["n_loc"]), "let_ep": int(wl["n_let_ep"]), "let_sp": int(wl["n_let_sp"]), "cells": int(len(wl["tree_cells"])),
            "groups": int(b.n_walk), "list_entries_ep": int(prof["n_epj"] / steps), "list_entries_sp": int(prof["n_spj"] / steps),
            "interactions_ep": int(prof["n_interaction_ep"] / steps), "interactions_sp": int(prof["n_interaction_sp"] / steps),
            "device_bytes_used_by_the_step": int(free0 - free1), "nccl_bytes_per_step": int(st.nccl_bytes_per_step),
            "h2d_bytes_per_step": int(prof["h2d_bytes"] / steps), "d2h_bytes_per_step": int(prof["d2h_bytes"] / steps),
            "ms_per_step": sec * 1e3, "timeline_ms": tl,
            "spot_check_kernel_coords0": {"groups": [int(x) for x in gs], "n": int(len(ea)), "acc_median": float(np.median(ea)), "acc_max": float(ea.max()),
                                          "pot_max": float(ep.max()), "n_ngb_mismatches": nbad}}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        inter = sum(r["interactions_ep"] + r["interactions_sp"] for r in allr)
        ms = max(r["ms_per_step"] for r in allr)
        ok = all(r["spot_check_kernel_coords0"]["acc_median"] <= 1e-6 and r["spot_check_kernel_coords0"]["acc_max"] <= 1e-4 and
                 r["spot_check_kernel_coords0"]["n_ngb_mismatches"] == 0 for r in allr)
        print("CONFIG5 " + json.dumps({
            "config": {"workload": f"plummer_kroupa_N{n_star}_bin{int(round(100 * f_bin))}pct_artificial", "n_particles": n_star, "n_tree_particles": n_tree,
                       "n_gpus": world, "theta": 0.3, "r_out": prm["r_out"], "parallelism": f"domain_decomposition_x{world}",
                       "api": "device-resident tree step (pb_tree_upload_let, pb_upload_j_range, NCCL LET, pb_tree_force_resident)"},
            "ms_per_step": ms, "ginteractions_per_s": inter / (ms * 1e-3) * 1e-9, "interactions_per_step": inter,
            "device_gb_used_max": max(r["device_bytes_used_by_the_step"] for r in allr) / 1e9, "device_gb_total": total / 1e9,
            "host_seconds": {"generate_particles": t_gen, "build_domain_workload_total": t_build},
            "parity_spot_check_pass": ok, "ranks": allr}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
