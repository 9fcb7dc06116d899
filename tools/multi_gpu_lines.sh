#!/bin/bash
# Builder-run multi-GPU lines of the three bench workloads on N GPUs of one box:  bash tools/multi_gpu_lines.sh N TAG
N=$1; TAG=$2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$T bench.py --gpus $N --workload c5_scan --steps 3 --warmup 1 2> gpurun_out/${TAG}_scan_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_scan_${N}gpu.json
$T bench.py --gpus $N --workload c4_train --steps 5 --warmup 3 2> gpurun_out/${TAG}_train_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_train_${N}gpu.json
$T bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/${TAG}_bench_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu.json
$T bench.py --gpus $N --workload c3_bin_4view_512x640 --steps 40 --warmup 5 2> gpurun_out/${TAG}_c3_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_c3_${N}gpu.json
python - <<P
import json
for f in ["scan","train","bench","c3"]:
    try:
        l=json.load(open("gpurun_out/${TAG}_%s_${N}gpu.json"%f)); print(f, "value", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), l.get("collective",{}).get("allreduce_ms") if f=="train" else "")
    except Exception as e: print(f, "ERR", e)
P
