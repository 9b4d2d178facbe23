#!/bin/bash
# Multi-GPU validation (run with gpurun --gpus 2): sharded == unsharded for every transport, exchange timings, the sharded
# bench workloads.
tag=${1:-r2g}
G=${2:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py > gpurun_out/${tag}_dist_check.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_dist_check.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $G > gpurun_out/${tag}_bench_c3_${G}gpu.json 2> gpurun_out/${tag}_bench_c3_${G}gpu.log; echo "rc=$?" >> gpurun_out/${tag}_bench_c3_${G}gpu.log
for tr in allgather all_to_all p2p_barrier; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $G --transport $tr --steps 12 --no-extras > gpurun_out/${tag}_bench_c3_${G}gpu_$tr.json 2> gpurun_out/${tag}_bench_c3_${G}gpu_$tr.log; echo "rc=$?" >> gpurun_out/${tag}_bench_c3_${G}gpu_$tr.log
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $G --workload c4 > gpurun_out/${tag}_bench_c4_${G}gpu.json 2> gpurun_out/${tag}_bench_c4_${G}gpu.log; echo "rc=$?" >> gpurun_out/${tag}_bench_c4_${G}gpu.log
grep -E "PASS|FAIL|exchange of|rc=" gpurun_out/${tag}_dist_check.log | tail -30
tail -n 4 gpurun_out/${tag}_bench_*.log
