"""Extract the judged metrics of an .ncu-rep into a small CSV (committed under profiles/):
    python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/r1_xxx.csv"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    cols = [k for k in KEEP if k in idx]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols + ["top_stalls"])
        w.writerow([units[idx[c]] for c in cols] + ["share of sampled stalls"])
        for r in rows[2:]:
            vals = sorted(((float(r[idx[s]] or 0), s.replace("smsp__pcsamp_warps_issue_stalled_", "")) for s in stall), reverse=True)
            tot = sum(v for v, _ in vals) or 1.0
            top = " ".join(f"{n}={100 * v / tot:.0f}%" for v, n in vals[:5])
            w.writerow([r[idx[c]] for c in cols] + [top])
    print(f"{out}: {len(rows) - 2} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
