#!/bin/bash
# Round-2 session 5, GPU call 4: packed fp32 arithmetic (FADD2 / FMUL2 / FFMA2) in the register-resident phase kernels: gate + A/B.
tag=${1:-r5d}
mkdir -p gpurun_out
VP=$PWD/dps_ttc_b200/build_variants/libdpsttc_packed.so
DPSTTC_LIB=$VP timeout 200 python tools/phase_reg_check.py --n 3 > gpurun_out/${tag}_phase_reg_check_packed.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_phase_reg_check_packed.log
DPSTTC_LIB=$VP timeout 200 python -m pytest tests -m gpu -q -k "phase" > gpurun_out/${tag}_pytest_phase_packed.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_phase_packed.log
for rep in 1 2; do for v in scalar packed; do for n in 32 8; do
  lib=""; [ $v == packed ] && lib=$VP
  DPSTTC_LIB=$lib timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph --only phasefused > gpurun_out/${tag}_kb_${v}_n${n}_$rep.jsonl 2> gpurun_out/${tag}_kb_${v}_n${n}_$rep.err
  echo "$v n=$n rep=$rep: $(cut -c100-200 gpurun_out/${tag}_kb_${v}_n${n}_$rep.jsonl)"
done; done; done
DPSTTC_LIB=$VP timeout 120 ncu --set full --clock-control none --import-source on -k "regex:phase_" -s 9 -c 3 -f -o gpurun_out/${tag}_packed_n32 \
  python tools/kernel_bench.py --n 32 --only phasefused --iters 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -n 3 gpurun_out/${tag}_phase_reg_check_packed.log gpurun_out/${tag}_pytest_phase_packed.log gpurun_out/${tag}_ncu.log
grep -E "FAIL| g:| r:" gpurun_out/${tag}_phase_reg_check_packed.log | grep " 11 " | head
