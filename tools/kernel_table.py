"""Markdown table from tools/kernel_bench.py outputs:
    python tools/kernel_table.py profiles/r1_kernel_bench.jsonl profiles/r1_kernel_bench_n8.jsonl > profiles/r1_kernel_bench.md"""
import json
import sys


def load(path):
    out = {}
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            d = json.loads(line)
            out[d["kernel"]] = d
    return out


def main(big_path, small_path):
    big, small = load(big_path), load(small_path)
    peak = next(iter(big.values()))["peak_gbs"]
    print(f"# Graft kernels — micro-benchmark (tools/kernel_bench.py), measured HBM peak {peak} GB/s\n")
    print("HBM regime: N = 128 particles per launch (32 for phase retrieval), rotating buffers larger than L2, CUDA events "
          "around back-to-back launches.  Launch regime: N = 8 (bench.py's workload), launches replayed from one CUDA graph "
          "over argument sets that together exceed 2x L2.\n")
    print("| kernel (algorithmic bytes) | N | µs | GB/s | % of peak | N=8 µs | N=8 GB/s | N=8 % of peak |")
    print("|---|---|---|---|---|---|---|---|")
    for name, d in big.items():
        s = small.get(name)
        row = f"| {name} | {d['n_particles']} | {d['mean_us']:.1f} | {d['gbs']:.0f} | {100 * d['frac_of_measured_peak']:.1f} |"
        row += f" {s['mean_us']:.2f} | {s['gbs']:.0f} | {100 * s['frac_of_measured_peak']:.1f} |" if s else " – | – | – |"
        print(row)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
