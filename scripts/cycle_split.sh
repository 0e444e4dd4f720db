#!/bin/bash
# Per-phase split of the fixed-point cycle of `solve_biomath input/biocase.prm 6 6` (C++ driver, one GPU) and the
# launch list of its kernels.  Usage: gpurun -- bash scripts/cycle_split.sh [tag]
tag=${1:-r2g}
mkdir -p gpurun_out /tmp/cyc && cd /tmp/cyc
export MSGPU_DRIVER_TIMING=2 MSGPU_CYCLE_LIMIT=6 MSGPU_SOLUTION_FILES=0
$GRAFT_REPO_ROOT/dealiiscale_b200/_build/solve_biomath $GRAFT_REPO_ROOT/input/biocase.prm 6 6 > $GRAFT_REPO_ROOT/gpurun_out/${tag}_cycle_split.log 2>&1
echo "driver rc=$?"
MSGPU_CYCLE_LIMIT=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/${tag}_cycle_launches.csv \
  $GRAFT_REPO_ROOT/dealiiscale_b200/_build/solve_biomath $GRAFT_REPO_ROOT/input/biocase.prm 6 6 > $GRAFT_REPO_ROOT/gpurun_out/${tag}_cycle_ncu.log 2>&1
echo "ncu rc=$?"
grep -E "timing" $GRAFT_REPO_ROOT/gpurun_out/${tag}_cycle_split.log | tail -24
