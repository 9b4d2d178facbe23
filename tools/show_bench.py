"""Pretty-print a kernel_bench jsonl file: python tools/show_bench.py gpurun_out/kernel_bench_*.jsonl"""
import json
import sys

for path in sys.argv[1:]:
    print(f"== {path}")
    for line in open(path):
        try:
            d = json.loads(line)
        except Exception:  # noqa: BLE001
            print(line[:200].rstrip())
            continue
        print(f"{d['kernel']:42s} N={d['n_particles']:4d} {d['mean_us']:9.2f} us {d['gbs']:8.1f} GB/s  "
              f"{100 * d['frac_of_measured_peak']:5.1f}% of measured peak")
