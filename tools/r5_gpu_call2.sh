#!/bin/bash
# Round-2 session 5, GPU call 2: gate + A/B of the register-resident ROW kernels of the phase guidance (2 vs 3 CTAs/SM builds),
# on top of the register column kernel (1-D grid, light column group last).
#   gpurun --timeout 420 -- 'bash tools/r5_gpu_call2.sh r5b'
tag=${1:-r5b}
mkdir -p gpurun_out
V3=dps_ttc_b200/build_variants/libdpsttc_rows3.so
timeout 200 python tools/phase_reg_check.py --n 3 > gpurun_out/${tag}_phase_reg_check.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_phase_reg_check.log
DPSTTC_LIB=$PWD/$V3 timeout 200 python tools/phase_reg_check.py --n 3 > gpurun_out/${tag}_phase_reg_check_rows3.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_phase_reg_check_rows3.log
run() {  # name cols rows lib
  for n in 32 8; do
    DPSTTC_PHASE_COLS_REG=$2 DPSTTC_PHASE_ROWS_REG=$3 DPSTTC_LIB=$4 timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph --only phasefused \
      > gpurun_out/${tag}_kb_$1_n$n.jsonl 2> gpurun_out/${tag}_kb_$1_n$n.err
    echo "$1 n=$n: $(cut -c100-260 gpurun_out/${tag}_kb_$1_n$n.jsonl)"
  done
}
run c1r0 1 0 ""
run c1r1 1 1 ""
run c1r1x3 1 1 $PWD/$V3
run c0r0 0 0 ""
# per-kernel times (ncu launch lists; numbers under ncu are not bench values) and one full capture of the default-candidate build
list() {  # name lib
  DPSTTC_PHASE_COLS_REG=1 DPSTTC_PHASE_ROWS_REG=1 DPSTTC_LIB=$2 timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:phase_" -s 9 -c 6 --csv \
    --log-file gpurun_out/${tag}_list_$1_n32.csv python tools/kernel_bench.py --n 32 --only phasefused --iters 3 > /dev/null 2>&1
}
list c1r1 ""
list c1r1x3 $PWD/$V3
DPSTTC_PHASE_COLS_REG=1 DPSTTC_PHASE_ROWS_REG=1 timeout 120 ncu --set full --clock-control none --import-source on -k "regex:phase_" -s 9 -c 3 -f -o gpurun_out/${tag}_phasefused_n32 \
  python tools/kernel_bench.py --n 32 --only phasefused --iters 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -n 2 gpurun_out/${tag}_phase_reg_check.log gpurun_out/${tag}_phase_reg_check_rows3.log gpurun_out/${tag}_ncu.log
grep -h "gpu__time_duration" gpurun_out/${tag}_list_*_n32.csv | cut -d, -f5,12- | head -20
