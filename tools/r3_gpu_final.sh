#!/bin/bash
# final 1-GPU regression of a round: tests, smoke, every workload's bench line, kernel tables, ncu launch list + full capture of the bench kernels
tag=${1:-r3s}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 300 python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c2.log
for w in c1 c3 c4; do
  timeout 250 python bench.py --workload $w --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_$w.log
done
timeout 300 python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c5.log
for n in 8 32; do timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph > gpurun_out/${tag}_kb_n$n.jsonl 2> gpurun_out/${tag}_kb_n$n.err; done
timeout 90 python tools/kernel_bench.py --n 128 > gpurun_out/${tag}_kb_n128.jsonl 2> gpurun_out/${tag}_kb_n128.err
# ncu passes come after the un-profiled runs above exited; numbers printed under ncu are never bench values
K='regex:resize_|posterior_update|particle_norms|norm_coef'
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv \
  --log-file gpurun_out/${tag}_bench_graft_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --eager-unet > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 -c 4 -f -o gpurun_out/${tag}_bench_kernels_n8 \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --eager-unet > gpurun_out/${tag}_ncu_full.log 2>&1
tail -n 8 gpurun_out/${tag}_pytest.log; tail -n 2 gpurun_out/${tag}_smoke.log gpurun_out/${tag}_bench_*.log gpurun_out/${tag}_ncu_list.log gpurun_out/${tag}_ncu_full.log
