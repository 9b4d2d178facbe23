#!/bin/bash
# One ncu --set full capture of one kernel_bench selector: tools/r2_ncu_one.sh <tag> <only> <n> [kernel-name regex]
tag=$1; only=$2; n=${3:-128}; rx=${4:-.}
mkdir -p gpurun_out
timeout 200 python tools/kernel_bench.py --n $n --only $only > gpurun_out/${tag}_kb.jsonl 2> gpurun_out/${tag}_kb.err || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s 3 -c 1 -f -o gpurun_out/${tag} \
  python tools/kernel_bench.py --n $n --only $only --iters 3 > gpurun_out/${tag}_ncu.log 2>&1
cat gpurun_out/${tag}_kb.jsonl; tail -3 gpurun_out/${tag}_ncu.log
