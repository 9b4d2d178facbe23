#!/bin/bash
# One GPU call that re-validates the tree: GPU suite, smoke(), the default bench line, kernel tables.
#   gpurun --timeout 600 -- 'bash tools/gpu_round_check.sh <tag>'   →  gpurun_out/<tag>_*
tag=${1:-check}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 240 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_1gpu.log
for n in 8 32; do timeout 120 python tools/kernel_bench.py --n $n --iters 50 --graph > gpurun_out/${tag}_kb_n$n.jsonl 2> gpurun_out/${tag}_kb_n$n.err; done
timeout 120 python tools/kernel_bench.py --n 128 > gpurun_out/${tag}_kb_n128.jsonl 2> gpurun_out/${tag}_kb_n128.err
tail -2 gpurun_out/${tag}_pytest.log gpurun_out/${tag}_smoke.log gpurun_out/${tag}_bench_1gpu.log
# A/B of the strip forward's TMA ring depth at small grids (default picks 6 stages up to 2 CTAs per SM)
for st in 3 6; do for n in 4 8 12; do
  DPSTTC_RESIZE_FWD_STAGES=$st timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph --only sr4 > gpurun_out/${tag}_ring${st}_n$n.jsonl 2> gpurun_out/${tag}_ring${st}_n$n.err
done; done
cat gpurun_out/${tag}_ring*_n*.jsonl | cut -c1-160
