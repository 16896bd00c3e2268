"""Rebuild the tracked round-2 ncu summaries under profiles/ from the scratch files of tools/gpu_r2_final.sh
(gpurun_out/z_*.ncu-rep, z_launches.csv, z_*.log) — run here, where ncu can read the reports."""
import csv, json, os, subprocess, sys
from collections import defaultdict, Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_static',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg.per_second', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__pcsamp_sample_count']


def raw_table(rep, out_csv):
    raw = subprocess.run(["ncu", "-i", os.path.join(O, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        print("no data in", rep); return None
    hdr, units = rows[0], rows[1]
    with open(os.path.join(P, out_csv), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch_{k + 1}" for k in range(len(rows) - 2)])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
    return hdr, units, rows[2:]


def col(t, name, scale={"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}):
    hdr, units, rows = t
    i = hdr.index(name)
    return [float(r[i]) * scale.get(units[i], 1.0) for r in rows]


def stalls(rep):
    """share of warp samples per stall reason + share of samples inside the packed-FP loops, from the source page"""
    src = subprocess.run(["ncu", "-i", os.path.join(O, rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"hdr": None, "rows": []}; blocks.append(cur); continue
        if cur is None: continue
        if cur["hdr"] is None: cur["hdr"] = r; continue
        cur["rows"].append(r)
    out = []
    for b in blocks:
        h = b["hdr"]; iS = h.index("# Samples"); iSrc = h.index("Source")
        sc = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        tot = sum(int(r[iS]) for r in b["rows"]) or 1
        agg = Counter()
        for r in b["rows"]:
            for i, n in sc: agg[n[6:]] += int(r[i] or 0)
        packed = sum(int(r[iS]) for r in b["rows"] if any(op in r[iSrc] for op in ("FFMA2", "FMUL2", "FADD2", "MUFU.RSQ", "FMNMX", "LDS.128", "LDS.64", "FSET", "FSEL")))
        out.append({"samples": tot, "stall_share_pct": {k: round(100 * v / tot, 1) for k, v in agg.most_common() if v},
                    "samples_on_pair_loop_instructions_pct": round(100 * packed / tot, 1)})
    return out


summary = ["# Round-2 ncu summary (B200, `--clock-control none`; per-launch times under ncu are cold-cache and serialised)\n"]
t = raw_table("z_prof_force.ncu-rep", "r2_ncu_force_kernel_N1e6.csv")
if t:
    rd, wr = col(t, "dram__bytes_read.sum"), col(t, "dram__bytes_write.sum")
    traffic = {"kernel": "pb::force_kernel<0,2,false,false,true>",
               "source": f"ncu --set full --clock-control none, {len(rd)} launches of `bench.py --steps 1 --warmup 1 --no-device-walk` (default workload: N=1e6 Kroupa Plummer, "
                         "10% binaries; 8 streams, so one launch = one of the 8 sub-batches of a 200-walk dispatch); summary in profiles/r2_ncu_force_kernel_N1e6.csv",
               "dram_bytes_per_launch": int((sum(rd) + sum(wr)) / len(rd)), "dram_bytes_read_per_launch": int(sum(rd) / len(rd)), "dram_bytes_write_per_launch": int(sum(wr) / len(wr)),
               "note": "cold-L2 figure of an isolated, profiled launch: the j store (71 MB for the whole step) and the index lists of the launch's walks are fetched from DRAM "
                       "again for every profiled launch, partial sums stay in L2. The algorithmic input of one launch is ~1.7 MB (0.76 GB H2D per step / 448 launches), so the "
                       "profiled traffic is ~3x that; at < 1 % of DRAM throughput it has no bearing on the bound."}
    json.dump(traffic, open(os.path.join(P, "r2_force_kernel_traffic.json"), "w"), indent=1)
    st = stalls("z_prof_force.ncu-rep")
    summary.append("## pb::force_kernel (functor path, one launch per stream and dispatch)\n")
    for name in ("gpu__time_duration.sum", "launch__registers_per_thread", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                 "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active"):
        i = t[0].index(name); summary.append(f"* `{name}` [{t[1][i]}]: " + ", ".join(r[i] for r in t[2]))
    summary.append(f"* DRAM per launch: {traffic['dram_bytes_per_launch'] / 1e6:.2f} MB")
    for k, s in enumerate(st): summary.append(f"* launch {k + 1}: warp-sample shares {s['stall_share_pct']}; {s['samples_on_pair_loop_instructions_pct']} % of samples on pair-loop instructions")
    summary.append("")
t = raw_table("z_prof_ws.ncu-rep", "r2_ncu_force_kernel_ws_N1e6.csv")
if t:
    st = stalls("z_prof_ws.ncu-rep")
    summary.append("## pb::force_kernel_ws (device-resident step: ONE persistent launch per tree step, N = 1e6)\n")
    for name in ("Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                 "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
                 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"):
        if name in t[0]:
            i = t[0].index(name); summary.append(f"* `{name}` [{t[1][i]}]: " + ", ".join(r[i] for r in t[2]))
    for s in st: summary.append(f"* warp-sample shares {s['stall_share_pct']}; {s['samples_on_pair_loop_instructions_pct']} % of samples on pair-loop instructions")
    summary.append("")
t = raw_table("z_prof_walk.ncu-rep", "r2_ncu_walk_kernel_N1e6.csv")
if t:
    summary.append("## pb::walk_kernel_c (compact records, single speculative pass, N = 1e6)\n")
    for name in ("Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum"):
        if name in t[0]:
            i = t[0].index(name); summary.append(f"* `{name}` [{t[1][i]}]: " + ", ".join(r[i] for r in t[2]))
    summary.append("")

src = os.path.join(O, "z_launches.csv")
if os.path.exists(src):
    txt = open(src).read()
    lr = [r for r in csv.reader(txt.splitlines()) if len(r) > 5]
    ik, iv = lr[0].index("Kernel Name"), lr[0].index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in lr[1:]:
        try: v = float(r[iv].replace(",", ""))
        except ValueError: continue
        n = r[ik].split("(")[0][:70]
        agg[n][0] += 1; agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    summary.append("## Launch list of `python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity` (ncu --metrics gpu__time_duration.sum, first 3000 launches)\n")
    summary.append("The command runs the recording functor step, then per warm-up / timed iteration one replay of the recorded step and its force-only timing pass, one functor step "
                   "(448 force launches each: the reduction is fused into the force kernel, no `reduce_kernel` launches are left) and one device-resident tree step (compact, walk, i-prep, plan, emit "
                   "and ONE persistent force launch that also reduces and writes the forces to host memory); the capture ends after 3000 launches, inside the second iteration.  In one tree step of "
                   "the product path the force kernel is > 99 % of the kernel time either way (`value`: 448 launches, 26.2 ms; `e2e`: one launch, 25.3 ms beside 2.6 ms of walk).\n")
    summary.append("| kernel | launches | total ms (under ncu) | share |\n|---|---|---|---|")
    for n, (c, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        summary.append(f"| `{n}` | {c} | {tt / 1e6:.3f} | {100 * tt / tot:.1f} % |")
    summary.append("")
    # keep the list itself, thinned to one line per launch
    with open(os.path.join(P, "r2_ncu_launches_N1e6.csv"), "w") as f:
        f.write(txt)
open(os.path.join(P, "r2_ncu_force_kernels.md"), "w").write("\n".join(summary) + "\n")
san = []
for name in ("z_memcheck.log", "z_racecheck.log"):
    p = os.path.join(O, name)
    if os.path.exists(p):
        lines = open(p).read().strip().splitlines()
        san.append(f"== compute-sanitizer {'memcheck' if 'mem' in name else 'racecheck'}: python tools/sanitize_small.py (functor path, neighbour search with pair emission, "
                   "host-planned and device-resident tree steps incl. the warp-specialised kernel, field query, changeover correction)\n" + "\n".join(lines[-4:]))
if san:
    open(os.path.join(P, "r2_compute_sanitizer.txt"), "w").write("\n\n".join(san) + "\n")
print("\n".join(summary)[:6000])
