#!/bin/bash
# Final GPU call of a round: validation (tools/gpu_round_check.sh minus the ring A/B) + ncu evidence for the bench's kernels.
#   gpurun --timeout 420 -- 'bash tools/gpu_round_final.sh <tag>'   →  gpurun_out/<tag>_*
tag=${1:-final}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 150 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_1gpu.log
for n in 8 32; do timeout 60 python tools/kernel_bench.py --n $n --iters 50 --graph > gpurun_out/${tag}_kb_n$n.jsonl 2> gpurun_out/${tag}_kb_n$n.err; done
timeout 60 python tools/kernel_bench.py --n 128 > gpurun_out/${tag}_kb_n128.jsonl 2> gpurun_out/${tag}_kb_n128.err
# ncu passes come after the un-profiled runs above exited; numbers printed under ncu are never bench values
K='regex:resize_|posterior_update|particle_norms|norm_coef'
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv \
  --log-file gpurun_out/${tag}_bench_graft_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --eager-unet > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 -c 8 -f -o gpurun_out/${tag}_bench_kernels_n8 \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --eager-unet > gpurun_out/${tag}_ncu_full.log 2>&1
tail -n 2 gpurun_out/${tag}_pytest.log gpurun_out/${tag}_smoke.log gpurun_out/${tag}_bench_1gpu.log gpurun_out/${tag}_ncu_list.log gpurun_out/${tag}_ncu_full.log
ls -la gpurun_out | tail -n 12
