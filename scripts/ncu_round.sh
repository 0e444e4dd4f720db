#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, then one --set full capture of the
# two heavy kernels on a reduced number of systems (ncu replays each kernel ~40 times).
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --macro-ref 4"
$CMD2 > gpurun_out/ncu_plain2.json 2> gpurun_out/ncu_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:'k_cg_strip|k_cg_resident|k_cell_moments|k_gather' -s 6 -c 3 -o gpurun_out/prof_$TAG $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
