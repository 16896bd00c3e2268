"""Collects the round-2 bench lines from gpurun_out/ into profiles/r2_bench_summary.md (verbatim JSON + a digest)."""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(R, "gpurun_out")

def line_of(fn):
    p = os.path.join(G, fn)
    if not os.path.exists(p): return None
    for line in open(p):
        if line.startswith('{"metric"') or line.startswith('{"impl"'):
            return line.strip()
    return None

RUNS = [
    ("1 GPU, default (`coords = 2`), 16-core host", "z_bench.log", "python bench.py"),
    ("1 GPU, `coords = 0` (kernel-level mode: walk-relative two-float dx for every pair)", "z_bench_coords0.log", "python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --opt coords=0"),
    ("1 GPU, reference arm", "z_bench_ref.log", "python bench.py --impl reference --steps 3 --warmup 1"),
    ("2 GPUs", "s_bench_2gpu.log", "torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 8 --warmup 3"),
    ("4 GPUs", "s_bench_4gpu.log", "torchrun --nproc-per-node 4 bench.py --gpus 4 --steps 8 --warmup 3"),
    ("8 GPUs, `raw_upload = 1` (default of the multi-rank stepper)", "s_bench_8gpu.log", "torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 8 --warmup 3"),
    ("8 GPUs, `raw_upload = 0` (one commit earlier: the functor path's kernel without the fused reduction)", "g8_bench_raw0.log", "PETAR_B200_RAW_UPLOAD=0 torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 8 --warmup 3 --no-parity"),
    ("1 GPU, BASELINE config 4 stand-in: N = 1e6 stars, 100 % binaries -> 7e6 tree particles", "c4_bench.log", "python bench.py --f-bin 1.0 --steps 3 --warmup 3 --no-cpu-baseline"),
    ("earlier in the round (before `sp2i`, fused reduction): 1 GPU, functor path with EP lists as plain indices (`ep_runs = 0`, the default)", "l_bench_runs0.log", "python bench.py --no-cpu-baseline --no-parity --opt ep_runs=0"),
    ("earlier in the round: 1 GPU, functor path with EP lists as runs (`ep_runs = 1`)", "l_bench_runs1.log", "python bench.py --no-cpu-baseline --no-parity --opt ep_runs=1"),
]

def digest(d):
    if d.get("impl") == "reference":
        return "reference arm: %.1f Ginteractions/s, %.0f ms per step, %d cores" % (d["value"], d["ms_per_step"], d["cpu_baseline"]["cores"])
    e = d["e2e"]; f = d.get("e2e_functors", {})
    s = "value %.1f G/s (%.2f ms" % (d["value"], d["ms_per_step"])
    vb = d.get("value_breakdown", {})
    if vb.get("let_exchange_ms"): s += " = kernels %.2f + LET exchange %.2f" % (vb["kernels_only_ms"], vb["let_exchange_ms"])
    s += "), roofline %.3f nominal / %.3f measured peak; e2e (device-resident tree step) %.2f ms = %.0f G/s, H2D %.0f MB" % (
        d["roofline"]["frac"], d["roofline"]["measured_fp32_peak"]["frac"], e["ms_per_step"], e["value"], e["h2d_bytes_per_step"] / 1e6)
    if f: s += "; functor path %.2f ms, H2D %.0f MB" % (f["ms_per_step"], f["h2d_bytes_per_step"] / 1e6)
    tl = e.get("device_timeline_ms_max_over_ranks")
    if tl: s += "; device timeline [ms] " + ", ".join("%s %.2f" % (k, v) for k, v in tl.items())
    hp = e.get("rank0_host_phases_ms")
    if hp: s += "; rank-0 host phases [ms] " + ", ".join("%s %.2f" % (k, v) for k, v in hp.items())
    par = d.get("parity", {})
    if par: s += "; parity pass=%s (drop-in max %.1e, kernel max %.1e over all ranks)" % (par.get("all_ranks", {}).get("pass"), par.get("all_ranks", {}).get("dropin_acc_max", float("nan")), par.get("all_ranks", {}).get("kernel_acc_max", float("nan")))
    return s

def main():
    head = open(os.path.join(R, "tools", "r2_bench_summary_head.md")).read()
    out = [head]
    out.append("\n## Bench lines (verbatim JSON from `bench.py`, one per run; a number printed under ncu is never listed here)\n")
    for title, fn, cmd in RUNS:
        ln = line_of(fn)
        if ln is None:
            print("missing", fn, file=sys.stderr); continue
        d = json.loads(ln)
        out.append("### %s\n\n`%s` (`gpurun_out/%s`)\n\n%s\n\n```json\n%s\n```\n" % (title, cmd, fn, digest(d), ln))
    open(os.path.join(R, "profiles", "r2_bench_summary.md"), "w").write("\n".join(out))
    print("wrote profiles/r2_bench_summary.md")

if __name__ == "__main__":
    main()
