"""Parity check of the multi-GPU path on real GPUs (run under torchrun): forces of every rank's
domain through NCCL LET exchange + CUDA kernels vs the fp64 oracle over the same lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_w = int(os.environ.get("WORLD_SIZE", "1"))
if os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _w))
import numpy as np
import torch
import torch.distributed as dist
from petar_b200 import engine, harness as hz, multigpu
from oracle import binding as ob

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
L = engine.load()
engine.check(L.pb_init(rank, lr), "pb_init")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
mass, pos, vel = hz.make_plummer(n)
prm = hz.petar_auto_params(mass, vel)
r_in, r_out, rs = hz.particle_rout_rsearch(mass, vel, prm)
wl = multigpu.build_domain_workload(pos, mass, vel, rs, r_in, r_out, rank, world, dist)
wl["prm"] = prm
st = multigpu.DomainStepper(wl, rank, world, dist)
b = wl["batch"]
f = np.zeros(b.n_epi_total, dtype=engine.ForceSoft)
for _ in range(2):
    st.step(f)
ref = ob.walks_index(b, prm["eps"], prm["r_out"], prm["G"])
ea = np.linalg.norm(f["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
ep = np.abs((f["pot"] - ref["pot"]) / ref["pot"])
ok = np.median(ea) <= 1e-6 and ea.max() <= 1e-4 and np.median(ep) <= 1e-6 and ep.max() <= 1e-4 and np.array_equal(f["n_ngb"], ref["n_ngb"])
print(f"rank {rank}/{world} n_loc={wl['n_loc']} let_ep={wl['n_let_ep']} let_sp={wl['n_let_sp']} nccl_bytes={st.nccl_bytes_per_step} "
      f"acc med {np.median(ea):.2e} max {ea.max():.2e} pot max {ep.max():.2e} nngb_equal {np.array_equal(f['n_ngb'], ref['n_ngb'])} -> {'OK' if ok else 'FAIL'}", flush=True)
# the same step with the lists built on the GPU from the global (local + LET) tree
f2 = np.zeros_like(f)
for _ in range(2):
    st.step_device_walk(f2)
ea2 = np.linalg.norm(f2["acc"] - ref["acc"], axis=1) / np.linalg.norm(ref["acc"], axis=1)
ok2 = np.median(ea2) <= 1e-6 and ea2.max() <= 1e-4 and np.array_equal(f2["n_ngb"], ref["n_ngb"])
print(f"rank {rank}/{world} device walk over the LET tree: acc med {np.median(ea2):.2e} max {ea2.max():.2e} nngb_equal {np.array_equal(f2['n_ngb'], ref['n_ngb'])} -> {'OK' if ok2 else 'FAIL'}", flush=True)
ok = ok and ok2
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
