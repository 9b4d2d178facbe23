#!/bin/bash
# A/B of the cluster-barrier variants of blur_fused.cu: stress (race) check + N=128 / N=8 timing per variant library
tag=${1:-r2u}
mkdir -p gpurun_out
for v in base pd01 pd10 pd11; do
  lib=""; [ $v != base ] && lib="dps_ttc_b200/build_variants/libdpsttc_$v.so"
  echo "== $v" >> gpurun_out/${tag}_ab.log
  DPSTTC_LIB=$lib timeout 200 python tools/fused_stress.py 96 40 >> gpurun_out/${tag}_ab.log 2>&1
  DPSTTC_LIB=$lib timeout 200 python tools/kernel_bench.py --n 128 --only gaussfused 2>/dev/null | cut -c1-150 >> gpurun_out/${tag}_ab.log
  DPSTTC_LIB=$lib timeout 200 python tools/kernel_bench.py --n 8 --graph --only gaussfused 2>/dev/null | cut -c1-150 >> gpurun_out/${tag}_ab.log
done
cat gpurun_out/${tag}_ab.log
