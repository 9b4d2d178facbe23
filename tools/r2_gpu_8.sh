#!/bin/bash
# 8-GPU validation: the driver's scaling command (default workload under torchrun = c3), c4 strong-scaled, c5 at spec, dist_check
tag=${1:-r2m}
G=8
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 300 $T --master-port 29534 bench.py --gpus $G --steps 20 --warmup 3 > gpurun_out/${tag}_bench_c3_${G}gpu.json 2> gpurun_out/${tag}_bench_c3_${G}gpu.log; echo "c3 rc=$?"
timeout 300 $T --master-port 29536 bench.py --gpus $G --workload c4 > gpurun_out/${tag}_bench_c4_${G}gpu.json 2> gpurun_out/${tag}_bench_c4_${G}gpu.log; echo "c4 rc=$?"
timeout 400 $T --master-port 29537 bench.py --gpus $G --workload c5 --steps 12 --no-extras > gpurun_out/${tag}_bench_c5_${G}gpu.json 2> gpurun_out/${tag}_bench_c5_${G}gpu.log; echo "c5 rc=$?"
timeout 300 $T --master-port 29533 tools/dist_check.py > gpurun_out/${tag}_dist_check_${G}gpu.log 2>&1; echo "dist rc=$?"
grep -E "exchange of|FAIL" gpurun_out/${tag}_dist_check_${G}gpu.log | tail -8
python - <<PY
import json
for w in ("c3","c4","c5"):
    try:
        d=json.load(open(f"gpurun_out/${tag}_bench_{w}_8gpu.json"))
        print(w, round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d.get("sharded_bit_identical"), {k:v for k,v in d.get("exchange",{}).items() if k!="collectives"}, d.get("single_gpu_same_workload",{}).get("value"))
    except Exception as e:
        print(w, "ERR", e)
PY
