"""Per-source-line summary of an .ncu-rep captured with --import-source on (kernels compiled with -lineinfo):
stall samples, executed warp instructions, shared-memory wavefronts (ideal / excessive) per CUDA source line.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]"""
import csv
import re
import subprocess
import sys


def main(rep, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = None
    out = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            # "Source" appears twice (cuda, sass): first index is the CUDA text
            continue
        if hdr is None or len(r) < 10 or not r[0].strip().isdigit():
            continue
        def g(k):
            m = re.match(r"[-+0-9.eE]+", r[hdr[k]])
            return float(m.group(0)) if m and m.group(0) not in ("-", "+") else 0.0
        out.append((int(r[0]), r[1].strip()[:110], g("# Samples"), g("Instructions Executed"), g("L1 Wavefronts Shared"),
                    g("L1 Wavefronts Shared Excessive"), g("stall_short_sb"), g("stall_long_sb"), g("stall_barrier"), g("stall_wait"), g("stall_mio"), g("stall_math")))
    ts = sum(o[2] for o in out) or 1
    ti = sum(o[3] for o in out) or 1
    print(f"total samples {ts:.0f}  warp instructions {ti:.0f}")
    print("line  %samp  %inst  smem_wf  excess  short long barrier wait mio math | source")
    for o in sorted(out, key=lambda o: -o[2])[:top]:
        print(f"{o[0]:4d} {100*o[2]/ts:6.1f} {100*o[3]/ti:6.1f} {o[4]:9.0f} {o[5]:8.0f} {o[6]:5.0f} {o[7]:5.0f} {o[8]:5.0f} {o[9]:5.0f} {o[10]:5.0f} {o[11]:5.0f} | {o[1]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
