#!/bin/bash
# generic A/B: tools/r2_gpu_ab.sh <tag> <only> "<variant names>" "<n list>"   (variant "base" = the in-tree library)
tag=$1; only=$2; vars=$3; ns=${4:-"8 128"}
mkdir -p gpurun_out
for v in $vars; do
  lib=""; [ $v != base ] && lib="dps_ttc_b200/build_variants/libdpsttc_$v.so"
  for n in $ns; do
    extra=""; [ $n -lt 100 ] && extra="--graph"
    echo -n "$v n=$n: " >> gpurun_out/${tag}_ab.log
    DPSTTC_LIB=$lib timeout 200 python tools/kernel_bench.py --n $n --only $only $extra 2>/dev/null | python -c "
import sys, json
print('; '.join(f\"{json.loads(l)['kernel'][:24]} {json.loads(l)['mean_us']}us {json.loads(l)['frac_of_measured_peak']}\" for l in sys.stdin if l.startswith('{')))" >> gpurun_out/${tag}_ab.log
  done
done
cat gpurun_out/${tag}_ab.log
