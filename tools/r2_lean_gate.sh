#!/bin/bash
# Gate + A/B of the opt-in lean kernels (SR strip forward, short-strip adjoint, streaming forward W pass, phase column epilogue) (written after round 1's last GPU call; see profiles/r1l_resize_n8_sass.md).
#   gpurun --timeout 420 -- 'bash tools/r2_lean_gate.sh r2a'   →  gpurun_out/<tag>_lean_*
# 1. bit-identity against the default kernels (must print PASS twice per op), 2. µs per launch at N = 4 / 8 / 12 / 32.
tag=${1:-r2a}
mkdir -p gpurun_out
for n in 3 8; do
  timeout 200 python tools/variant_check.py --op sr4 --n $n \
    --env DPSTTC_RESIZE_FWD_LEAN=0,DPSTTC_RESIZE_ADJ_LEAN=0 --env DPSTTC_RESIZE_FWD_LEAN=1,DPSTTC_RESIZE_ADJ_LEAN=1 \
    --env DPSTTC_RESIZE_FWD_LEAN=1,DPSTTC_RESIZE_FWD_STAGES=6,DPSTTC_RESIZE_ADJ_LEAN=1 > gpurun_out/${tag}_lean_gate_n$n.log 2>&1
  echo "gate n=$n rc=$?" >> gpurun_out/${tag}_lean_gate_n$n.log
done
for v in 0 1; do for n in 4 8 12 32; do
  DPSTTC_RESIZE_FWD_LEAN=$v DPSTTC_RESIZE_ADJ_LEAN=$v timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph --only sr4 \
    > gpurun_out/${tag}_lean${v}_n$n.jsonl 2> gpurun_out/${tag}_lean${v}_n$n.err
done; done
# streaming SR forward with the lean W pass (same flag; the streaming family is forced so that a small n reaches it)
timeout 200 python tools/variant_check.py --op sr4 --n 12 \
  --env DPSTTC_RESIZE_VARIANT=stream,DPSTTC_RESIZE_FWD_LEAN=0 --env DPSTTC_RESIZE_VARIANT=stream,DPSTTC_RESIZE_FWD_LEAN=1 \
  --env DPSTTC_RESIZE_VARIANT=big,DPSTTC_RESIZE_FWD_LEAN=1 > gpurun_out/${tag}_lean_gate_stream.log 2>&1
echo "gate stream rc=$?" >> gpurun_out/${tag}_lean_gate_stream.log
for v in 0 1; do
  DPSTTC_RESIZE_FWD_LEAN=$v DPSTTC_RESIZE_ADJ_LEAN=$v timeout 90 python tools/kernel_bench.py --n 128 --only sr4 > gpurun_out/${tag}_lean${v}_n128.jsonl 2> gpurun_out/${tag}_lean${v}_n128.err
done
# phase retrieval: lean output epilogue of the column kernel (DPSTTC_PHASE_LEAN=1)
timeout 200 python tools/variant_check.py --op phase --n 4 --env DPSTTC_PHASE_LEAN=0 --env DPSTTC_PHASE_LEAN=1 > gpurun_out/${tag}_lean_gate_phase.log 2>&1
echo "gate phase rc=$?" >> gpurun_out/${tag}_lean_gate_phase.log
for v in 0 1; do for n in 8 32; do
  DPSTTC_PHASE_LEAN=$v timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph --only phase > gpurun_out/${tag}_phase_lean${v}_n$n.jsonl 2> gpurun_out/${tag}_phase_lean${v}_n$n.err
done; done
tail -n 3 gpurun_out/${tag}_lean_gate_n*.log gpurun_out/${tag}_lean_gate_stream.log gpurun_out/${tag}_lean_gate_phase.log; cat gpurun_out/${tag}_lean*_n*.jsonl | cut -c1-150
