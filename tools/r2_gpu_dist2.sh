#!/bin/bash
tag=${1:-r2k}
G=${2:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 500 $T --master-port 29534 bench.py --gpus $G > gpurun_out/${tag}_bench_c3_${G}gpu.json 2> gpurun_out/${tag}_bench_c3_${G}gpu.log; echo "rc=$?" >> gpurun_out/${tag}_bench_c3_${G}gpu.log
timeout 500 $T --master-port 29536 bench.py --gpus $G --workload c4 > gpurun_out/${tag}_bench_c4_${G}gpu.json 2> gpurun_out/${tag}_bench_c4_${G}gpu.log; echo "rc=$?" >> gpurun_out/${tag}_bench_c4_${G}gpu.log
timeout 600 $T --master-port 29537 bench.py --gpus $G --workload c5 --steps 12 > gpurun_out/${tag}_bench_c5_${G}gpu.json 2> gpurun_out/${tag}_bench_c5_${G}gpu.log; echo "rc=$?" >> gpurun_out/${tag}_bench_c5_${G}gpu.log
timeout 300 $T --master-port 29538 bench.py --gpus $G --workload c2 --no-extras > gpurun_out/${tag}_bench_c2_${G}gpu.json 2> gpurun_out/${tag}_bench_c2_${G}gpu.log; echo "rc=$?" >> gpurun_out/${tag}_bench_c2_${G}gpu.log
timeout 300 $T --master-port 29539 bench.py --gpus $G --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref_${G}gpu.json 2> gpurun_out/${tag}_bench_ref_${G}gpu.log; echo "rc=$?" >> gpurun_out/${tag}_bench_ref_${G}gpu.log
if [ "$G" -ge 4 ]; then
timeout 400 $T --master-port 29533 tools/dist_check.py > gpurun_out/${tag}_dist_check_${G}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_dist_check_${G}gpu.log
grep -E "exchange of|FAIL|rc=" gpurun_out/${tag}_dist_check_${G}gpu.log | tail -8
fi
tail -n 3 gpurun_out/${tag}_bench_*_${G}gpu.log
