#!/bin/bash
# Round 2, GPU call 1: lean-kernel gates (bit identity + A/B), the GPU suite with the new pins, the c2 and c3 bench lines.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${tag}_gpus.txt
for n in 3 8; do
  timeout 200 python tools/variant_check.py --op sr4 --n $n \
    --env DPSTTC_RESIZE_FWD_LEAN=0,DPSTTC_RESIZE_ADJ_LEAN=0 --env DPSTTC_RESIZE_FWD_LEAN=1,DPSTTC_RESIZE_ADJ_LEAN=1 \
    --env DPSTTC_RESIZE_FWD_LEAN=1,DPSTTC_RESIZE_FWD_STAGES=6,DPSTTC_RESIZE_ADJ_LEAN=1 > gpurun_out/${tag}_lean_gate_n$n.log 2>&1
  echo "gate n=$n rc=$?" >> gpurun_out/${tag}_lean_gate_n$n.log
done
timeout 200 python tools/variant_check.py --op sr4 --n 12 \
  --env DPSTTC_RESIZE_VARIANT=stream,DPSTTC_RESIZE_FWD_LEAN=0 --env DPSTTC_RESIZE_VARIANT=stream,DPSTTC_RESIZE_FWD_LEAN=1 \
  --env DPSTTC_RESIZE_VARIANT=big,DPSTTC_RESIZE_FWD_LEAN=1 > gpurun_out/${tag}_lean_gate_stream.log 2>&1
echo "gate stream rc=$?" >> gpurun_out/${tag}_lean_gate_stream.log
timeout 200 python tools/variant_check.py --op phase --n 4 --env DPSTTC_PHASE_LEAN=0 --env DPSTTC_PHASE_LEAN=1 > gpurun_out/${tag}_lean_gate_phase.log 2>&1
echo "gate phase rc=$?" >> gpurun_out/${tag}_lean_gate_phase.log
for v in 0 1; do for n in 8 32; do
  DPSTTC_RESIZE_FWD_LEAN=$v DPSTTC_RESIZE_ADJ_LEAN=$v DPSTTC_PHASE_LEAN=$v timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph \
    > gpurun_out/${tag}_lean${v}_n$n.jsonl 2> gpurun_out/${tag}_lean${v}_n$n.err
done
DPSTTC_RESIZE_FWD_LEAN=$v DPSTTC_RESIZE_ADJ_LEAN=$v DPSTTC_PHASE_LEAN=$v timeout 90 python tools/kernel_bench.py --n 128 > gpurun_out/${tag}_lean${v}_n128.jsonl 2> gpurun_out/${tag}_lean${v}_n128.err
done
timeout 420 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
DPSTTC_RESIZE_FWD_LEAN=1 DPSTTC_RESIZE_ADJ_LEAN=1 DPSTTC_PHASE_LEAN=1 timeout 300 python -m pytest tests -m gpu -q \
  -k "super or resolution or sr or variants or c2 or phase or c4 or kernels or dropin or graphed" > gpurun_out/${tag}_pytest_lean.log 2>&1; echo "pytest-lean rc=$?" >> gpurun_out/${tag}_pytest_lean.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 300 python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c2.log
timeout 200 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c3.log
tail -n 3 gpurun_out/${tag}_lean_gate_*.log
tail -n 25 gpurun_out/${tag}_pytest.log; tail -n 5 gpurun_out/${tag}_pytest_lean.log gpurun_out/${tag}_smoke.log gpurun_out/${tag}_bench_c2.log gpurun_out/${tag}_bench_c3.log
