#!/usr/bin/env python
"""bench.py — soft-force hot path benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 1000000] [--impl reference]

One "step" = one pass of the hot path over one tree step's worth of walks of the workload
(N-particle Plummer model, theta 0.3, n_group_limit 512, 200 walks per dispatch — the shape FDPS
hands to the dispatch functor at reference src/petar.hpp:894-899):

* value  — interactions/s with all packed inputs resident in HBM, device-timed.  N = 1: every recorded dispatch's
           force + reduce kernels (pb_replay).  N > 1: per step the LET exchange as well — EP rows gathered on the
           device, SP rows copied, one NCCL all-to-all per kind into the peers' j stores, publication — then the
           same kernels; the kernels-only figure stays as a sub-key;
* e2e    — from HOST buffers, wall clock, the SAME API at every N: the device-resident tree step (C ABI:
           pb_tree_upload[_let] + pb_upload_j[_range] + pb_tree_force_resident) — the tree and this rank's own
           particles go up, (N > 1: LET over NCCL,) interaction lists, task plan and forces are made on the GPU,
           forces come down.  It does MORE than the reference arm it is compared with (which is handed its lists);
* e2e_functors — the drop-in step through the PeTar functor boundary (C++ shim -> C ABI) from host buffers, lists given:
           j upload, per-dispatch packing, H2D, kernels, D2H, scatter into ForceSoft arrays; reported at every N;
* roofline — force-kernel time vs the non-tensor FP32 peak with the north-star flop convention
           (38 flop per EP-EP, 65 per EP-SP interaction);
* cpu_baseline — the reference's own AVX-512/AVX2 kernels (oracle/_ref) on this box's host cores.

Multi-GPU (torchrun, one rank per GPU): the particles are split into N spatial domains, each rank
owns one, the local-essential-tree j (EP near, SP far) travel rank-to-rank in the device j format
through one NCCL all-to-all per step straight into the receivers' j stores; forces need no
reduction.  Total work is fixed as N grows ("strong" scaling).  Every line carries a parity report per rank sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the host-side packing is OpenMP code, so give
# each rank its share of the host cores (must happen before any OpenMP runtime is loaded)
_world = int(os.environ.get("WORLD_SIZE", "1"))
if _world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    _ref_arm = "reference" in sys.argv          # the reference arm runs on rank 0 alone: it gets every host core
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // (1 if _ref_arm else _world)))

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_EP, FLOP_SP = 38.0, 65.0          # north-star convention (BASELINE.json)
N_SM, FP32_LANES = 148, 128


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per force-kernel launch from the committed ncu
    --set full capture (it cannot be measured inside a timed run); None if no capture is committed."""
    for name in ("r2_force_kernel_traffic.json", "r1_force_kernel_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)["dram_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            continue
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  Started before the warm-up (the
    tool needs a few hundred ms to deliver its first sample) at 20 ms intervals; every sample carries the host
    time it arrived at, and only those inside [t_begin, t_end] of the timed region are used.  A timed region
    too short to catch 3 samples (multi-GPU runs last tens of ms) falls back to the samples of warm-up + timed
    region together — the same kernels, back to back — and says so."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def stats(rows):
            sm, mx, reasons, power = [], [], set(), []
            for _, r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                    for k, nm in enumerate(names):
                        if r[5 + k].lower().startswith("active"):
                            reasons.add(nm)
                except (ValueError, IndexError):
                    continue
            # under load = samples in the upper half of the observed clock range
            load = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
            return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                    "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}

        inside = [x for x in self.rows if t_begin is None or (t_begin <= x[0] <= (t_end or 1e300))]
        out = stats(inside)
        out["window"] = "timed region"
        if out["samples"] < 3:
            out = stats(self.rows)
            out["window"] = ("warm-up + timed region (the timed region, %.0f ms, is shorter than 3 sampling intervals)"
                             % (1e3 * ((t_end or 0) - (t_begin or 0))))
        return out


def build_workload(n, rank, world, args):
    """Particles, parameters and this rank's walk batch (+ LET plan for world > 1)."""
    from petar_b200 import harness as hz
    t0 = time.time()
    if args.workload == "plummer":
        mass, pos, vel = hz.make_plummer(n)
        prm = hz.petar_auto_params(mass, vel)
        P = hz.plummer_particles(mass, pos, vel, prm)
        r_in, r_out, rs, ptype = P["r_in"], P["r_out"], P["rs"], None
        wl = {"prm": prm, "n": n, "n_tree": n, "name": f"plummer_equal_mass_N{n}"}
    else:
        # BASELINE.json configs[2]: Kroupa IMF, 10 % of the stars in binaries, artificial particles
        P = hz.kroupa_binary_particles(n, f_bin=args.f_bin, seed=1)
        mass, pos, vel, rs, r_in, r_out, ptype, prm = (P[k] for k in ("mass", "pos", "vel", "rs", "r_in", "r_out", "ptype", "prm"))
        wl = {"prm": prm, "n": n, "n_tree": len(mass), "n_bin": P["n_bin"],
              "name": f"plummer_kroupa_N{n}_bin{int(round(100 * args.f_bin))}pct_artificial"}
    wl["P"] = P
    if world == 1:
        batch, epi_src = hz.build_walk_batch(pos, mass, rs, vel=vel, r_in=r_in, r_out=r_out, ptype=ptype)
        wl["batch"], wl["epi_src"], wl["let"] = batch, epi_src, None
    else:
        from petar_b200 import multigpu
        wl.update(multigpu.build_domain_workload(pos, mass, vel, rs, r_in, r_out, rank, world, ptype=ptype))
    wl["t_build"] = time.time() - t0
    return wl


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU SIMD kernels (oracle/_ref) over the same walk lists,
    all host threads, rank 0 only."""
    if rank != 0:
        return
    from oracle import binding as ob
    wl = build_workload(args.n, 0, 1, args)
    batch, prm = wl["batch"], wl["prm"]
    isa = "avx512" if ob._cpu_has_avx512() else "avx2"
    kind = "reference" if ob.ref_available(isa) else "port"
    cores = os.cpu_count()
    # bounded sample: as many leading walks as keep one step near `--cpu-seconds`
    I_ep, I_sp = batch.interactions()
    nw = batch.n_walk
    force = ob.new_force(batch.n_epi_total)

    def one(nwalks):
        sl = slice(0, nwalks)
        if kind == "reference":
            _, sec = ob.ref_walks_index(batch, prm["eps"], prm["r_out"], prm["G"], isa=isa, n_threads=0, walk_slice=sl, force=force)
        else:
            t0 = time.perf_counter()
            ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"], walk_slice=sl)
            sec = time.perf_counter() - t0
        return sec

    probe_w = max(1, min(nw, 64))
    sec = one(probe_w)
    ie, isp = batch.interactions(slice(0, probe_w))
    rate = (ie + isp) / sec
    target = args.cpu_seconds * rate
    cum = np.cumsum(batch.n_epi.astype(np.int64) * (batch.n_epj.astype(np.int64) + batch.n_spj))
    nwalks = int(min(nw, max(1, np.searchsorted(cum, target) + 1)))
    ie, isp = batch.interactions(slice(0, nwalks))
    for _ in range(args.warmup):
        one(min(nwalks, probe_w))
    times = [one(nwalks) for _ in range(args.steps)]
    sec = float(np.mean(times))
    gint = (ie + isp) / sec * 1e-9
    sample = f"first {nwalks} of {nw} walks of one tree step ({ie + isp:.3e} interactions per step), {isa}, OpenMP dynamic over walks"
    line = {
        "impl": "reference", "metric": "soft-force Ginteractions/s", "value": gint, "unit": "Ginteractions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * (I_ep + I_sp) / (ie + isp),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, 1),
        "cpu_baseline": {"value": gint, "unit": "Ginteractions/s", "cores": cores, "kind": kind, "cpu_model": cpu_model(), "sample": sample},
        "e2e": {"value": gint, "unit": "Ginteractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sec_per_nbody_time_unit": sec * (I_ep + I_sp) / (ie + isp) / prm["dt_soft"],
    }
    print(json.dumps(line), flush=True)


def workload_config(args, wl, world):
    prm = wl["prm"]
    return {"workload": wl["name"], "n_particles": args.n, "n_tree_particles": wl["n_tree"], "theta": 0.3, "n_leaf_limit": 20,
            "n_group_limit": 512, "n_walk_limit": args.n_walk_limit, "r_out": prm["r_out"], "dt_soft": prm["dt_soft"], "eps": prm["eps"],
            "multipole": "quadrupole", "parallelism": f"domain_decomposition_x{world}",
            "l2_policy": "inputs larger than L2: every step re-reads all dispatches' index lists, i-particles and the j store"}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def parity_report(wl, f_default, world, rank, run_sample_coords0):
    """SURVEY §8(d) "parity report (every run, every rank)" on a bounded sample of this rank's walks (first, middle and
    last 32) against the fp64 oracle.  Two statements (checker only, oracle/):
      * drop-in: the default-mode (coords = 2) forces of the timed e2e step + the float-replay changeover correction an
        unmodified PeTar applies, against oracle forces + all-double correction (oracle/dropin_check.py);
      * kernel: the same walks re-run with coords = 0 against the NoSimd functors directly."""
    from oracle import binding as ob
    from oracle.dropin_check import DropinChecker
    from petar_b200 import harness as hz
    batch, prm, P = wl["batch"], wl["prm"], wl["P"]
    nw = batch.n_walk
    starts = sorted({0, max(0, nw // 2 - 16), max(0, nw - 32)})
    walks = np.unique(np.concatenate([np.arange(w0, min(nw, w0 + 32)) for w0 in starts]))
    rows = np.concatenate([np.arange(batch.i_off[w], batch.i_off[w + 1]) for w in walks])
    sub = type(batch)(batch.epj, batch.spj, batch.epi[rows], np.concatenate([[0], np.cumsum(batch.n_epi[walks])]),
                      np.concatenate([batch.id_epj[batch.ej_off[w]:batch.ej_off[w + 1]] for w in walks]), np.concatenate([[0], np.cumsum(batch.n_epj[walks])]),
                      np.concatenate([batch.id_spj[batch.sj_off[w]:batch.sj_off[w + 1]] for w in walks]), np.concatenate([[0], np.cumsum(batch.n_spj[walks])]))
    ref = ob.walks_index(sub, prm["eps"], prm["r_out"], prm["G"])
    f0 = run_sample_coords0(sub)

    def stats(f, r):
        ea = np.linalg.norm(f["acc"] - r["acc"], axis=1) / np.maximum(np.linalg.norm(r["acc"], axis=1), 1e-300)
        ep = np.abs(f["pot"] - r["pot"]) / np.maximum(np.abs(r["pot"]), 1e-300)
        return {"acc_rel_err": {"median": float(np.median(ea)), "p99": float(np.percentile(ea, 99)), "max": float(ea.max())},
                "pot_rel_err": {"median": float(np.median(ep)), "p99": float(np.percentile(ep, 99)), "max": float(ep.max())},
                "n_ngb_mismatches": int((f["n_ngb"] != r["n_ngb"]).sum()), "n": int(len(f))}

    kern = stats(f0, ref)
    kern["pass"] = bool(kern["acc_rel_err"]["median"] <= 1e-6 and kern["acc_rel_err"]["max"] <= 1e-4 and
                        kern["pot_rel_err"]["median"] <= 1e-6 and kern["pot_rel_err"]["max"] <= 1e-4 and kern["n_ngb_mismatches"] == 0)
    # the particle set this rank's store holds (own particles, then the LET particles): correction inputs in store order
    p0_all = hz.corr_particles(P)
    if world == 1:
        chk = DropinChecker(P, prm, subset=wl["epi_src"][rows])
    else:
        gid = wl["store_gid"]
        chk = DropinChecker(None, prm, subset=wl["epi_src"][rows], p0=p0_all[gid], rs=P["rs"][gid])
    drop = chk.compare(f_default[rows], ref)
    drop["n_ngb_mismatches"] = int((f_default["n_ngb"][rows] != ref["n_ngb"]).sum())
    drop["pass"] = bool(drop["pass"] and drop["n_ngb_mismatches"] == 0)
    return {"sample": f"{len(rows)} i-particles of {len(walks)} walks (first, middle, last) of rank {rank} against the fp64 oracle",
            "tolerance": "acc/pot relative error <= 1e-6 median, <= 1e-4 max; neighbour counts equal",
            "dropin_coords2_plus_float_replay": drop, "kernel_coords0": kern, "pass": bool(drop["pass"] and kern["pass"])}


def cpu_baseline_leg(args, wl):
    from oracle import binding as ob
    batch, prm = wl["batch"], wl["prm"]
    isa = "avx512" if ob._cpu_has_avx512() else "avx2"
    kind = "reference" if ob.ref_available(isa) else "port"
    nw = batch.n_walk
    force = ob.new_force(batch.n_epi_total)

    def run(nwalks):
        if kind == "reference":
            _, sec = ob.ref_walks_index(batch, prm["eps"], prm["r_out"], prm["G"], isa=isa, walk_slice=slice(0, nwalks), force=force)
        else:
            t0 = time.perf_counter(); ob.walks_index(batch, prm["eps"], prm["r_out"], prm["G"], walk_slice=slice(0, nwalks)); sec = time.perf_counter() - t0
        return sec

    probe_w = max(1, min(nw, 32))
    sec = run(probe_w)
    ie, isp = batch.interactions(slice(0, probe_w))
    rate = (ie + isp) / sec
    cum = np.cumsum(batch.n_epi.astype(np.int64) * (batch.n_epj.astype(np.int64) + batch.n_spj))
    nwalks = int(min(nw, max(1, np.searchsorted(cum, args.cpu_seconds * rate) + 1)))
    sec = run(nwalks)
    ie, isp = batch.interactions(slice(0, nwalks))
    return {"value": (ie + isp) / sec * 1e-9, "unit": "Ginteractions/s", "cores": os.cpu_count(), "kind": kind, "cpu_model": cpu_model(),
            "sample": f"first {nwalks} of {nw} walks of one tree step ({ie + isp:.3e} interactions, {sec:.2f} s), {isa}, OpenMP over walks"}


def fp32_peak():
    """The measured non-tensor FP32 peak of this pool's B200 (tools/microbench4.cu, profiles/r2_fp32_peak.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_fp32_peak.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=1000000, help="number of particles (BASELINE metric: 1e6)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work per step of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e", default="tree", choices=["tree", "functors"],
                    help="which step is the `e2e` key: the device-resident tree step (default, every N) or the drop-in functor path; "
                         "the other one is always reported next to it")
    ap.add_argument("--no-device-walk", action="store_true", help="skip the tree step leg altogether (implies --e2e functors; kernel tuning runs)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--nr", type=int, default=0)
    ap.add_argument("--cull", type=int, default=1)
    ap.add_argument("--jchunk", type=int, default=0)
    ap.add_argument("--occupancy", type=int, default=2)
    ap.add_argument("--workload", default="kroupa_binaries", choices=["kroupa_binaries", "plummer"],
                    help="kroupa_binaries = BASELINE.json configs[2] stand-in (default); plummer = equal-mass Plummer (configs[1] shape)")
    ap.add_argument("--f-bin", type=float, default=0.1, help="fraction of stars in binaries (kroupa_binaries); 1.0 = BASELINE configs[3]/[4] shape")
    ap.add_argument("--n-walk-limit", type=int, default=200, help="walks per dispatch; PeTar fixes 200 (src/petar.hpp:888)")
    ap.add_argument("--opt", action="append", default=[], help="extra library option key=value (pb_set_option), repeatable")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from petar_b200 import engine
    L = engine.load()
    engine.check(L.pb_init(rank, local_rank), "pb_init")
    for k, v in (("streams", args.streams), ("nr", args.nr), ("cull", args.cull), ("jchunk", args.jchunk), ("occupancy", args.occupancy)):
        engine.set_option(k, v)
    for kv in args.opt:
        k, v = kv.split("=")
        engine.set_option(k, int(v))
    coords_mode = engine.get_option("coords")

    wl = build_workload(args.n, rank, world, args)
    batch, prm = wl["batch"], wl["prm"]
    I_ep, I_sp = batch.interactions()
    eps, r_out, G = prm["eps"], prm["r_out"], prm["G"]
    force = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)
    force_dw = np.zeros(batch.n_epi_total, dtype=engine.ForceSoft)

    stepper = None
    if world > 1:
        from petar_b200 import multigpu
        stepper = multigpu.DomainStepper(wl, rank, world, dist)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # FDPS holds the per-group pointer tables ready when it calls dispatch; build them once
    tables = engine.make_dispatch_tables(batch, force, args.n_walk_limit)

    def functor_step():
        if stepper is None:
            engine.calc_force_all_and_write_back(batch, eps, r_out, G, force=force, my_rank=rank, tables=tables)
        else:
            stepper.step(force)

    if stepper is None:
        # the tree is written straight into the library's pinned staging buffers, as a converter from FDPS's cells would do it
        cells, groups = batch.tree.export_tree(out=engine.tree_stage(batch.tree.n_nodes, batch.n_walk))
        tree_step = lambda: engine.tree_force(batch, cells, groups, eps, r_out, G, force=force_dw, resident=True)
    else:
        cells, groups = wl["tree_cells"], wl["tree_groups"]
        tree_step = lambda: stepper.step_device_walk(force_dw)
    # the result array of the tree step persists from step to step (as FDPS's force array does): the library page-locks it
    # once and the force kernel writes the reduced forces straight into it
    engine.set_option("raw_result", int(os.environ.get("PETAR_B200_RAW_RESULT", "1")))
    run_tree = not args.no_device_walk
    if not run_tree:
        args.e2e = "functors"

    # ---- record one functor step so its packed inputs stay resident in HBM ----
    engine.check(L.pb_record_begin(), "pb_record_begin")
    functor_step()
    engine.check(L.pb_record_end(), "pb_record_end")
    launches_per_step = L.pb_replay_launches()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    ms_t, ms_f = C.c_float(0), C.c_float(0)

    def value_step():
        """One device-resident step; returns (ms of the LET exchange on the device, ms of all kernels, ms of the force kernels)."""
        ms_x = 0.0
        if stepper is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            stepper.let_exchange_only()
            ev1.record()
            ev1.synchronize()
            ms_x = ev0.elapsed_time(ev1)
        engine.check(L.pb_replay(1, C.byref(ms_t), C.byref(ms_f)), "pb_replay")
        return ms_x, float(ms_t.value), float(ms_f.value)

    # ---- warm-up ----
    for _ in range(args.warmup):
        value_step()
        functor_step()
        if run_tree:
            tree_step()

    # ---- timed: K steps, device resident, device-timed ----
    barrier()
    t_timed_begin = time.time()
    acc = np.zeros(3)
    for _ in range(args.steps):
        acc += value_step()
    barrier()
    ms_xchg, ms_kern, ms_force = (acc / args.steps).tolist()
    ms_step = ms_xchg + ms_kern

    # ---- timed: K steps end to end from host buffers, both APIs ----
    engine.get_profile(reset=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        functor_step()
    barrier()
    sec_fun = (time.perf_counter() - t0) / args.steps
    prof = engine.get_profile()
    sec_tree, prof_tree, timeline = 0.0, None, None
    if run_tree:
        engine.get_profile(reset=True)
        if stepper is not None:
            stepper.host_dw, stepper.n_dw = {}, 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            tree_step()
        barrier()
        sec_tree = (time.perf_counter() - t0) / args.steps
        prof_tree = engine.get_profile()
        timeline = engine.tree_timeline()
    t_timed_end = time.time()
    clock_probe_s = 0.0
    if t_timed_end - t_timed_begin < 0.5:
        # too short for nvidia-smi to sample (multi-GPU runs): keep the same recorded steps running, untimed,
        # for one more second so that the clocks are read under the very same load
        tp = time.time()
        while time.time() - tp < 1.0:
            engine.check(L.pb_replay(1, C.byref(ms_t), C.byref(ms_f)), "pb_replay")
        clock_probe_s = time.time() - tp
        t_timed_end = time.time()
    clocks = sampler.stop(t_timed_begin, t_timed_end) if rank == 0 else None
    if clocks is not None and clock_probe_s > 0:
        clocks["window"] += " + %.1f s of the same recorded steps replayed right after it (untimed)" % clock_probe_s

    # ---- parity of this rank's results (checker only; after the timed regions) ----
    f_default = (force_dw if args.e2e == "tree" else force).copy()          # the e2e leg's result
    parity = None
    if not args.no_parity:
        def run_sample_coords0(sub):
            engine.set_option("coords", 0)
            try:
                return engine.calc_force_all_and_write_back(sub, eps, r_out, G, my_rank=rank).copy()
            finally:
                engine.set_option("coords", coords_mode)
        try:
            parity = parity_report(wl, f_default, world, rank, run_sample_coords0)
            if run_tree:      # the other API's result must say the same: identical lists, another summation order
              parity["tree_step_vs_functors_max_rel_acc"] = float(np.abs(force_dw["acc"] - force["acc"]).max() / np.abs(force["acc"]).max())
              parity["tree_step_n_ngb_equal"] = bool(np.array_equal(force_dw["n_ngb"], force["n_ngb"]))
        except Exception as ex:  # noqa: BLE001 — the checker must never take the bench down
            parity = {"unavailable": repr(ex), "pass": False}

    # ---- max over ranks, totals over ranks ----
    vals = torch.tensor([ms_step, ms_force, sec_fun, sec_tree, ms_kern, ms_xchg] + ([timeline[k] for k in engine.TIMELINE_KEYS] if timeline else [0.0] * 6),
                        dtype=torch.float64, device="cuda")
    tot = torch.tensor([I_ep, I_sp, prof["h2d_bytes"] / args.steps, prof["d2h_bytes"] / args.steps,
                        stepper.nccl_bytes_per_step if stepper else 0,
                        (prof_tree["h2d_bytes"] / args.steps) if prof_tree else 0, (prof_tree["d2h_bytes"] / args.steps) if prof_tree else 0],
                       dtype=torch.float64, device="cuda")
    par = torch.tensor([1.0 if (parity and parity.get("pass")) else 0.0] +
                       ([parity["dropin_coords2_plus_float_replay"]["acc_rel_err"]["max"], parity["kernel_coords0"]["acc_rel_err"]["max"],
                         parity["dropin_coords2_plus_float_replay"]["acc_rel_err"]["median"], parity["kernel_coords0"]["acc_rel_err"]["median"]]
                        if parity and "kernel_coords0" in parity else [0.0] * 4), dtype=torch.float64, device="cuda")
    par_min = par.clone()
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(par, op=dist.ReduceOp.MAX)
        dist.all_reduce(par_min, op=dist.ReduceOp.MIN)
    ms_step, ms_force, sec_fun, sec_tree, ms_kern, ms_xchg = (float(x) for x in vals[:6].tolist())
    tl_max = dict(zip(engine.TIMELINE_KEYS, (float(x) for x in vals[6:].tolist())))
    I_ep_t, I_sp_t, h2d, d2h, nccl_b, h2d_tree, d2h_tree = (float(x) for x in tot.tolist())

    if rank == 0:
        peaks, peak_src = measured_peaks()
        f_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        peak_tf = N_SM * FP32_LANES * 2 * f_mhz * 1e6 / 1e12 * world
        flops = FLOP_EP * I_ep_t + FLOP_SP * I_sp_t
        ach_tf = flops / (ms_force * 1e-3) / 1e12
        inter = I_ep_t + I_sp_t
        fun = {"value": inter / sec_fun * 1e-9, "unit": "Ginteractions/s", "ms_per_step": sec_fun * 1e3,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "nccl_bytes_per_step": nccl_b,
               "rank0_ms_per_step": {"host_pack_unpack": prof["t_copy"] * 1e3 / args.steps,
                                     "host_plan": prof["t_plan"] * 1e3 / args.steps, "host_pack": prof["t_pack"] * 1e3 / args.steps,
                                     "host_unpack": prof["t_unpack"] * 1e3 / args.steps, "host_enqueue": prof["t_enqueue"] * 1e3 / args.steps, "h2d": prof["t_send"] * 1e3 / args.steps,
                                     "kernels": prof["t_calc"] * 1e3 / args.steps, "d2h": prof["t_recv"] * 1e3 / args.steps,
                                     "gpu_idle_between_walk_groups": prof["t_gap"] * 1e3 / args.steps,
                                     "note": "device intervals of concurrent streams overlap; they do not add up to ms_per_step"},
               "api": "CalcForceWithLinearCutoffCUDAMultiWalk / RetrieveForceCUDA driven by the FDPS-style walk-group loop (C++ shim -> C ABI), host buffers"}
        if stepper is not None and stepper.n_steps:
            fun["rank0_step_phases_ms"] = {k: v * 1e3 / stepper.n_steps for k, v in stepper.host_s.items()}
        tree = None
        if run_tree:
            tree = {"value": inter / sec_tree * 1e-9, "unit": "Ginteractions/s", "ms_per_step": sec_tree * 1e3,
                    "h2d_bytes_per_step": h2d_tree, "d2h_bytes_per_step": d2h_tree, "nccl_bytes_per_step": nccl_b,
                    "n_cells_rank0": int(len(cells)), "n_groups_rank0": int(len(groups)),
                    "device_timeline_ms_max_over_ranks": tl_max,
                    "host_walk_ms_for_context": batch.tree.timing()[1] * 1e3,
                    "api": "pb_tree_upload[_let] + pb_upload_j / pb_upload_j_range + (N > 1: pb_let_gather_epj, NCCL all-to-all, pb_publish_j) + "
                           "pb_tree_force_resident (C ABI), host buffers: the tree and this rank's particles go up, interaction lists, task plan "
                           "and forces are made on the GPU, forces come down; host_walk_ms = the harness's OpenMP walk that builds the lists "
                           "the functor path is handed for free"}
            if stepper is not None and stepper.n_dw:
                tree["rank0_host_phases_ms"] = {k: v * 1e3 / stepper.n_dw for k, v in stepper.host_dw.items()}
                if stepper.trace:
                    tree["rank0_let_enqueue_trace_ms"] = {k: v * 1e3 / stepper.trace["n"] for k, v in stepper.trace.items() if k != "n"}
        e2e, other, other_key = (dict(tree), fun, "e2e_functors") if args.e2e == "tree" else (dict(fun), tree, "e2e_tree_step")
        fp = fp32_peak()
        line = {
            "metric": "soft-force Ginteractions/s", "value": inter / (ms_step * 1e-3) * 1e-9, "unit": "Ginteractions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl, world),
            "value_breakdown": {"kernels_only_ms": ms_kern, "kernels_only_ginteractions_per_s": inter / (ms_kern * 1e-3) * 1e-9,
                                "let_exchange_ms": ms_xchg, "force_kernel_ms": ms_force,
                                "what": "device time, max over ranks: LET exchange (device gather of EP rows, NCCL all-to-all into the peers' j stores, "
                                        "publication; N > 1 only) + force and reduce kernels of every recorded dispatch, inputs resident in HBM"},
            "interactions_per_step": {"ep_ep": I_ep_t, "ep_sp": I_sp_t},
            "sec_per_nbody_time_unit": {"device_resident": ms_step * 1e-3 / prm["dt_soft"], "e2e_hot_path": e2e["ms_per_step"] * 1e-3 / prm["dt_soft"],
                                        "note": "hot path only; FDPS tree build/walk, hard integrator etc. are outside this path"},
            "e2e": e2e,
            "coords_mode": coords_mode,
            "gpu_launches": int((launches_per_step + 2 * (world > 1)) * (args.steps + args.warmup) + prof["n_kernel_launch"] +
                                (prof_tree["n_kernel_launch"] if prof_tree else 0)),
            "roofline": {"bound": "fp32", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                         "traffic": ncu_traffic(), "traffic_unit": "DRAM bytes per launch (ncu capture, profiles/)",
                         "kernel": "pb::force_kernel", "ms_per_step_kernel": ms_force,
                         "force_kernel_launches_per_step": launches_per_step // 2,
                         "algorithmic_flop_per_launch": flops / max(1, (launches_per_step // 2) * world),
                         "flop_convention": "38 per EP-EP, 65 per EP-SP interaction (north star)",
                         "peak_source": f"148 SM x 128 lanes x 2 x sm_max_mhz={f_mhz:.0f} from {peak_src} (nominal non-tensor FP32 lanes; "
                                        "MEASURED_PEAKS has no FP32-pipe figure)",
                         "measured_fp32_peak": ({"peak": fp["lanes_per_clk_sm"] * N_SM * 2 * f_mhz * 1e6 / 1e12 * world,
                                                 "frac": ach_tf / (fp["lanes_per_clk_sm"] * N_SM * 2 * f_mhz * 1e6 / 1e12 * world),
                                                 "source": "profiles/r2_fp32_peak.json: %.1f of 128 lanes per clock and SM sustained by FFMA2 with <= 2 "
                                                           "distinct register operands (tools/microbench4.cu) x sm_max_mhz" % fp["lanes_per_clk_sm"]} if fp else None),
                         # the two rooflines the contract names, for the record: neither binds this kernel
                         "hbm": {"algorithmic_bytes_per_step": h2d, "achieved": h2d / (ms_force * 1e-3) / 1e9,
                                 "peak": float(peaks.get("hbm_gbs", 6536.4)) * world, "unit": "GB/s",
                                 "frac": h2d / (ms_force * 1e-3) / 1e9 / (float(peaks.get("hbm_gbs", 6536.4)) * world),
                                 "note": "every byte the step's kernels must read once (tables, i-particles, index lists, j store = the H2D bytes)"},
                         "note": "bound is the non-tensor FP32/issue pipe, not HBM or tensor: ~300-500 flop per HBM byte"},
            "clocks": clocks,
        }
        if other is not None:
            line[other_key] = other
        if world > 1:
            line["e2e"]["omp_threads_per_rank"] = int(os.environ.get("OMP_NUM_THREADS", "0"))
        if parity is not None:
            parity["all_ranks"] = {"pass": bool(par_min[0].item() >= 0.5), "dropin_acc_max": float(par[1].item()), "kernel_acc_max": float(par[2].item()),
                                   "dropin_acc_median_max_over_ranks": float(par[3].item()), "kernel_acc_median_max_over_ranks": float(par[4].item())}
            parity["pass"] = parity["all_ranks"]["pass"]
            line["parity"] = parity
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_leg(args, wl)
            except Exception as ex:  # the checker must never take the bench down
                line["cpu_baseline"] = {"value": None, "unit": "Ginteractions/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}
            # informational second baseline: the reference's own CUDA kernels and host
            # protocol (blocking copies, default stream, 1 warp per block) on this same GPU and lists
            try:
                from oracle import binding as ob
                if ob.ref_cuda_available():
                    L.pb_finalize()                       # release our device buffers first
                    ob.ref_cuda_step(batch, eps, r_out, G)          # warm-up (allocations)
                    _, ms_k, ms_w = ob.ref_cuda_step(batch, eps, r_out, G)
                    line["reference_cuda_kernels"] = {
                        "kernels_ms_per_step": ms_k, "kernels_ginteractions_per_s": (I_ep + I_sp) / (ms_k * 1e-3) * 1e-9,
                        "e2e_ms_per_step": ms_w, "e2e_ginteractions_per_s": (I_ep + I_sp) / (ms_w * 1e-3) * 1e-9,
                        "note": "informational: reference src/force_gpu_cuda.cu device code compiled for sm_100a, host side restated "
                                "(oracle/ref_cuda_driver.cu); same GPU, same walk lists, 200 walks per dispatch"}
            except Exception as ex:  # noqa: BLE001
                line["reference_cuda_kernels"] = {"unavailable": repr(ex)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    L.pb_finalize()


if __name__ == "__main__":
    main()
