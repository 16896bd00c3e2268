"""Rebuild the tracked ncu summaries under profiles/ from the scratch files of `tools/gpu_final.sh`
(gpurun_out/f_prof_force.ncu-rep, f_launches.csv) — run here, where ncu can read the report."""
import csv, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", os.path.join(O, "f_prof_force.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_static',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg.per_second', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__pcsamp_sample_count']
with open(os.path.join(P, "r1_ncu_force_kernel_N1e6.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch_{k + 1}" for k in range(len(rows) - 2)])
    for k in keep:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in rows[2:]])


def col(name, unit_scale={"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}):
    i = hdr.index(name)
    return [float(r[i]) * unit_scale[units[i]] for r in rows[2:]]


rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
traffic = {
    "kernel": "pb::force_kernel<0,2>",
    "source": f"ncu --set full --clock-control none, {len(rd)} launches of `bench.py --steps 1 --warmup 1` (default workload: N=1e6 Kroupa Plummer, 10% binaries; "
              f"{n_streams} streams, so one launch = one of the {n_streams} sub-batches of a 200-walk dispatch); summary in profiles/r1_ncu_force_kernel_N1e6.csv",
    "dram_bytes_per_launch": int((sum(rd) + sum(wr)) / len(rd)),
    "dram_bytes_read_per_launch": int(sum(rd) / len(rd)),
    "dram_bytes_write_per_launch": int(sum(wr) / len(wr)),
    "note": "average of the captured launches (" + ", ".join(f"{x / 1e6:.2f} MB" for x in rd) + " read; partial sums and results stay in L2 until the reduce "
            "kernel / D2H). The algorithmic input of such a launch (index lists + i-particles of its walks, plus the j-store lines they touch) is of the same "
            "size, so DRAM traffic is within ~1.5x of the algorithmic bytes and irrelevant to the bound (< 1 % of DRAM throughput).",
}
json.dump(traffic, open(os.path.join(P, "r1_force_kernel_traffic.json"), "w"), indent=1)

src = os.path.join(O, "f_launches.csv")
open(os.path.join(P, "r1_ncu_launches_N1e6.csv"), "w").write(open(src).read())
lr = [r for r in csv.reader(open(src)) if len(r) > 5]
ik, iv = lr[0].index("Kernel Name"), lr[0].index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in lr[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    n = r[ik].split("(")[0][:60]
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:62s} {c:5d} launches {t / 1e6:9.3f} ms  {100 * t / tot:5.1f} %")
print(json.dumps(traffic)[:300])
