#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, then one --set full capture of the heavy kernels
# on a reduced number of systems (ncu replays each kernel ~40 times).  Each ncu pass only after the same command
# has exited 0 without ncu.  Usage: gpurun -- bash scripts/ncu_round.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-cycle"
$CMD > gpurun_out/ncu_plain_$TAG.json 2> gpurun_out/ncu_plain_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
# 17 x 17 macro points = 289 systems: one wave of the persistent CG kernel
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs --no-cycle --macro-ref 4"
$CMD2 > gpurun_out/ncu_plain2_$TAG.json 2> gpurun_out/ncu_plain2_$TAG.err && \
ncu --set full --clock-control none --import-source on -k regex:'k_cg_stencil|k_cell_moments|k_gather|k_face_moments' -s 8 -c 4 -o gpurun_out/prof_$TAG $CMD2 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
