#!/bin/bash
# full regression + refreshed tables:  gpurun --timeout 900 -- 'bash tools/r3_gpu_full.sh <tag>'
tag=${1:-r3a}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 300 python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c2.log
timeout 200 python bench.py --workload c1 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c1.log
for n in 8 32; do timeout 90 python tools/kernel_bench.py --n $n --iters 50 --graph > gpurun_out/${tag}_kb_n$n.jsonl 2> gpurun_out/${tag}_kb_n$n.err; done
timeout 90 python tools/kernel_bench.py --n 128 > gpurun_out/${tag}_kb_n128.jsonl 2> gpurun_out/${tag}_kb_n128.err
tail -n 12 gpurun_out/${tag}_pytest.log; tail -n 3 gpurun_out/${tag}_smoke.log; tail -n 4 gpurun_out/${tag}_bench_c2.log gpurun_out/${tag}_bench_c1.log
