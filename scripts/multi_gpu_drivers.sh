#!/bin/bash
# Two-rank runs of the C++ drivers against their single-process runs.  Usage: gpurun --gpus 2 -- bash scripts/multi_gpu_drivers.sh
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi_rank_drivers.py -x -q 2>&1 | tail -25 | tee gpurun_out/multi_rank_drivers.log
