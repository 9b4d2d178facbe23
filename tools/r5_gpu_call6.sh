#!/bin/bash
# Round-2 session 5, last experiment: PHASE_PACKED=1 (FADD2 for the complex adds only, scalar twiddle products): A/B first, then its gate.
tag=${1:-r5h}
mkdir -p gpurun_out
VH=$PWD/dps_ttc_b200/build_variants/libdpsttc_hybrid.so
for v in scalar hybrid; do
  lib=""; [ $v == hybrid ] && lib=$VH
  DPSTTC_LIB=$lib timeout 40 python tools/kernel_bench.py --n 32 --iters 50 --graph --only phase,phasefused > gpurun_out/${tag}_kb_${v}_n32.jsonl 2> gpurun_out/${tag}_kb_${v}_n32.err
done
cut -c1-170 gpurun_out/${tag}_kb_*_n32.jsonl
DPSTTC_LIB=$VH timeout 60 python tools/phase_reg_check.py --n 3 > gpurun_out/${tag}_phase_reg_check_hybrid.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_phase_reg_check_hybrid.log
tail -n 2 gpurun_out/${tag}_phase_reg_check_hybrid.log
DPSTTC_LIB=$VH timeout 60 python -m pytest tests -m gpu -q -k "phase" > gpurun_out/${tag}_pytest_phase_hybrid.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_phase_hybrid.log
tail -n 3 gpurun_out/${tag}_pytest_phase_hybrid.log
