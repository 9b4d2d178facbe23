#!/bin/bash
# fused SR guidance: parity tests + stress + micro-benchmark at three particle counts
tag=${1:-r2v}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "sr or deferred" > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 300 python tools/fused_stress.py 96 30 > gpurun_out/${tag}_stress.log 2>&1
for n in 8 32 128; do
  extra=""; [ $n != 128 ] && extra="--graph"
  timeout 200 python tools/kernel_bench.py --n $n --only sr4fused,sr8fused $extra > gpurun_out/${tag}_kb_n$n.jsonl 2> gpurun_out/${tag}_kb_n$n.err
done
tail -n 5 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_stress.log; cat gpurun_out/${tag}_kb_n*.jsonl | cut -c1-200
