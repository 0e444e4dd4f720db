#!/bin/bash
# bench.py on N GPUs of one box with the NVLink data counters of GPU 0 read before and after.
# Usage: gpurun --gpus N -- bash scripts/bench_multi_gpu.sh N [tag] [steps]
N=${1:-8}; tag=${2:-r2}; steps=${3:-10}
mkdir -p gpurun_out
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${tag}_nvlink_before.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps $steps --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
echo "bench rc=$?"
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${tag}_nvlink_after.txt 2>&1
tail -c 600 gpurun_out/${tag}_bench_${N}gpu.err
python - <<P
import json
d=json.loads(open("gpurun_out/${tag}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","exchange_check")}, d["e2e"]["ms_per_step"])
print(d.get("strong")); print(d["cycle"].get("ms_cycle"), d["cycle"].get("ms_micro"), d["cycle"].get("ms_macro"), d["cycle"].get("ms_stop_test"))
P
