#!/bin/bash
# Round-2 session 5, last GPU call: the committed tree as it is (all phase paths on the register kernels by default).
tag=${1:-r5g}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 45 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 25 python tools/kernel_bench.py --n 32 --iters 50 --graph --only phase,phasefused > gpurun_out/${tag}_kb_phase_n32.jsonl 2> gpurun_out/${tag}_kb_phase_n32.err
tail -n 3 gpurun_out/${tag}_pytest.log; tail -n 2 gpurun_out/${tag}_smoke.log; cut -c1-170 gpurun_out/${tag}_kb_phase_n32.jsonl
