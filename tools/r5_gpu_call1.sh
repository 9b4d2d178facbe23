#!/bin/bash
# Round-2 session 5, GPU call 1: gate + A/B of the register-resident phase column kernel, then the full suite on the tree.
#   gpurun --timeout 480 -- 'bash tools/r5_gpu_call1.sh r5a'
tag=${1:-r5a}
mkdir -p gpurun_out
timeout 150 python tools/phase_reg_check.py --n 3 > gpurun_out/${tag}_phase_reg_check.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_phase_reg_check.log
for reg in 1 0; do for n in 32 8; do
  DPSTTC_PHASE_COLS_REG=$reg timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph --only phasefused > gpurun_out/${tag}_kb_phase_reg${reg}_n$n.jsonl 2> gpurun_out/${tag}_kb_phase_reg${reg}_n$n.err
done; done
cat gpurun_out/${tag}_kb_phase_reg*_n*.jsonl | cut -c1-200
DPSTTC_PHASE_COLS_REG=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
# ncu after the un-profiled runs: launch list of the three guidance kernels, then one full capture of the column kernel
DPSTTC_PHASE_COLS_REG=1 timeout 120 ncu --set full --clock-control none --import-source on -k "regex:phase_" -s 9 -c 3 -f -o gpurun_out/${tag}_phasefused_n32 \
  python tools/kernel_bench.py --n 32 --only phasefused --iters 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -n 3 gpurun_out/${tag}_phase_reg_check.log gpurun_out/${tag}_pytest.log gpurun_out/${tag}_smoke.log gpurun_out/${tag}_ncu.log
