#!/bin/bash
tag=${1:-r2i}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 300 python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c2.log
timeout 300 python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c5.log
timeout 300 python bench.py --workload c4 --steps 12 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c4.log
timeout 300 python bench.py --workload c1 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c1.log
tail -n 14 gpurun_out/${tag}_pytest.log; tail -n 3 gpurun_out/${tag}_smoke.log; tail -n 6 gpurun_out/${tag}_bench_c2.log gpurun_out/${tag}_bench_c5.log gpurun_out/${tag}_bench_c4.log gpurun_out/${tag}_bench_c1.log
