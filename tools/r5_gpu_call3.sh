#!/bin/bash
# Round-2 session 5, GPU call 3: register column kernel compiled for 2 vs 3 CTAs per SM (row kernels at 3), gate of the 3-CTA build.
tag=${1:-r5c}
mkdir -p gpurun_out
V3=$PWD/dps_ttc_b200/build_variants/libdpsttc_cols3.so
export DPSTTC_PHASE_COLS_REG=1 DPSTTC_PHASE_ROWS_REG=1
DPSTTC_LIB=$V3 timeout 200 python tools/phase_reg_check.py --n 3 > gpurun_out/${tag}_phase_reg_check_cols3.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_phase_reg_check_cols3.log
DPSTTC_LIB=$V3 timeout 200 python -m pytest tests -m gpu -q -k "phase" > gpurun_out/${tag}_pytest_phase_cols3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_phase_cols3.log
for rep in 1 2; do for v in cols2 cols3; do for n in 32 8; do
  lib=""; [ $v == cols3 ] && lib=$V3
  DPSTTC_LIB=$lib timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph --only phasefused > gpurun_out/${tag}_kb_${v}_n${n}_$rep.jsonl 2> gpurun_out/${tag}_kb_${v}_n${n}_$rep.err
  echo "$v n=$n rep=$rep: $(cut -c100-200 gpurun_out/${tag}_kb_${v}_n${n}_$rep.jsonl)"
done; done; done
DPSTTC_LIB=$V3 timeout 120 ncu --set full --clock-control none --import-source on -k "regex:phase_cols" -s 3 -c 1 -f -o gpurun_out/${tag}_cols3_n32 \
  python tools/kernel_bench.py --n 32 --only phasefused --iters 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -n 3 gpurun_out/${tag}_phase_reg_check_cols3.log gpurun_out/${tag}_pytest_phase_cols3.log gpurun_out/${tag}_ncu.log
