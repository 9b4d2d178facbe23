#!/bin/bash
# Round-2 session 5, GPU call 5: gate + A/B of the register kernels for the ADJOINT of the two-kernel phase path (the fused path's
# last kernel shares the changed epilogue, so the phase tests run again).
tag=${1:-r5f}
mkdir -p gpurun_out
timeout 200 python tools/phase_reg_check.py --n 3 > gpurun_out/${tag}_phase_reg_check.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_phase_reg_check.log
DPSTTC_PHASE_ADJ_REG=1 timeout 200 python -m pytest tests -m gpu -q -k "phase or dropin or pins or edge" > gpurun_out/${tag}_pytest_phase_adjreg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_phase_adjreg.log
for reg in 0 1; do
  DPSTTC_PHASE_ADJ_REG=$reg timeout 60 python tools/kernel_bench.py --n 32 --iters 50 --graph --only phase,phasefused > gpurun_out/${tag}_kb_phase_adjreg${reg}_n32.jsonl 2> gpurun_out/${tag}_kb_phase_adjreg${reg}_n32.err
  DPSTTC_PHASE_ADJ_REG=$reg timeout 60 python tools/kernel_bench.py --n 8 --iters 50 --graph --only phase,phasefused > gpurun_out/${tag}_kb_phase_adjreg${reg}_n8.jsonl 2> gpurun_out/${tag}_kb_phase_adjreg${reg}_n8.err
done
cut -c1-160 gpurun_out/${tag}_kb_phase_adjreg*_n*.jsonl
tail -n 3 gpurun_out/${tag}_phase_reg_check.log gpurun_out/${tag}_pytest_phase_adjreg.log
grep -c " ok" gpurun_out/${tag}_phase_reg_check.log; grep "FAIL" gpurun_out/${tag}_phase_reg_check.log | head
