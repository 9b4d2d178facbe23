#!/bin/bash
# Round-2 session 5, final GPU call: the whole GPU suite with the register forward kernels switched on (everything else is the
# tree's default), smoke, the default bench line (c2), the phase workload (c4), kernel tables, A/B of the two-kernel phase forward.
#   gpurun --timeout 560 -- 'bash tools/r5_gpu_final.sh r5e'
tag=${1:-r5e}
mkdir -p gpurun_out
DPSTTC_PHASE_FWD_REG=1 timeout 260 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 200 python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c2.log
for reg in 0 1; do
  DPSTTC_PHASE_FWD_REG=$reg timeout 60 python tools/kernel_bench.py --n 32 --iters 50 --graph --only phase > gpurun_out/${tag}_kb_phase_fwdreg${reg}_n32.jsonl 2> gpurun_out/${tag}_kb_phase_fwdreg${reg}_n32.err
done
cut -c1-200 gpurun_out/${tag}_kb_phase_fwdreg*_n32.jsonl
timeout 120 python bench.py --workload c4 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.log; echo "bench rc=$?" >> gpurun_out/${tag}_bench_c4.log
for n in 32 8; do timeout 70 python tools/kernel_bench.py --n $n --iters 50 --graph > gpurun_out/${tag}_kb_n$n.jsonl 2> gpurun_out/${tag}_kb_n$n.err; done
timeout 70 python tools/kernel_bench.py --n 128 > gpurun_out/${tag}_kb_n128.jsonl 2> gpurun_out/${tag}_kb_n128.err
tail -n 4 gpurun_out/${tag}_pytest.log; tail -n 2 gpurun_out/${tag}_smoke.log gpurun_out/${tag}_bench_c2.log gpurun_out/${tag}_bench_c4.log
