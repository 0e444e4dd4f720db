#!/bin/bash
# configs[2] through the C++ driver: solve_biomath biocase.prm 6 6 (4225 macro points x 4225 micro DoFs), the macro
# points sharded over N processes / GPUs, against the single-process run.  Usage: gpurun --gpus N -- bash scripts/bio_multi_gpu.sh N
N=${1:-8}
mkdir -p gpurun_out results
export MSGPU_CYCLE_LIMIT=${MSGPU_CYCLE_LIMIT:-6} MSGPU_DRIVER_TIMING=1 MSGPU_SOLUTION_FILES=0
B=dealiiscale_b200/_build
( time $B/solve_biomath input/biocase.prm 6 6 ) > gpurun_out/bio66_1gpu.log 2>&1
( time python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    $B/solve_biomath input/biocase.prm 6 6 ) > gpurun_out/bio66_${N}gpu.log 2>&1
echo "rc=$?"
grep -E "residual|timing|Results|^[0-9]+, |real" gpurun_out/bio66_1gpu.log | tail -30
echo ----
grep -E "residual|timing|Results|^[0-9]+, |real" gpurun_out/bio66_${N}gpu.log | tail -30
echo ---- stop histories identical:
diff <(grep -E "residual" gpurun_out/bio66_1gpu.log) <(grep -E "residual" gpurun_out/bio66_${N}gpu.log) && echo yes
