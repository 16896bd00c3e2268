"""Print a one-line summary of bench.py JSON lines found in the given log files."""
import json
import sys

for f in sys.argv[1:]:
    try:
        lines = [x for x in open(f) if x.startswith("{")]
        d = json.loads(lines[-1])
        if d.get("impl") == "reference":
            print(f"{f}: REFERENCE {d['value']:.1f} G/s {d['ms_per_step']:.1f} ms/step cores {d['cpu_baseline']['cores']} {d['cpu_baseline']['kind']}")
            continue
        r, e = d["roofline"], d["e2e"]
        print(f"{f}: gpus {d['n_gpus']} value {d['value']:.1f} G/s step {d['ms_per_step']:.2f} ms kernel {r['ms_per_step_kernel']:.2f} ms "
              f"frac {r['frac']:.3f} | e2e {e['value']:.1f} G/s {e['ms_per_step']:.2f} ms h2d {e['h2d_bytes_per_step'] / 1e6:.0f} MB "
              f"nccl {e.get('nccl_bytes_per_step', 0) / 1e6:.1f} MB | launches {d['gpu_launches']} clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}"
              + (f" | cpu {d['cpu_baseline']['value']:.1f} G/s x{d['cpu_baseline']['cores']}" if d.get("cpu_baseline") and d["cpu_baseline"].get("value") else ""))
    except Exception as ex:  # noqa: BLE001
        print(f"{f}: ERR {ex!r}")
        try:
            print(open(f).read()[-1200:])
        except OSError:
            pass
