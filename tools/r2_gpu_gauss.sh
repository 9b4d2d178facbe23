#!/bin/bash
# fused separable-blur guidance: parity test + micro-benchmark at three particle counts
tag=${1:-r2q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "separable" > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
for n in 8 32 128; do
  extra=""; [ $n != 128 ] && extra="--graph"
  timeout 200 python tools/kernel_bench.py --n $n --only gaussfused $extra > gpurun_out/${tag}_kb_n$n.jsonl 2> gpurun_out/${tag}_kb_n$n.err
done
timeout 200 python tools/fused_stress.py 96 30 > gpurun_out/${tag}_stress.log 2>&1
if [ -f dps_ttc_b200/build_variants/libdpsttc_trace.so ]; then DPSTTC_LIB=dps_ttc_b200/build_variants/libdpsttc_trace.so timeout 100 python tools/sepf_trace.py 128 > gpurun_out/${tag}_trace.log 2>&1; fi
tail -n 5 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_stress.log gpurun_out/${tag}_trace.log; cat gpurun_out/${tag}_kb_n*.jsonl
