#!/bin/bash
# ncu evidence for the bench command: (1) launch list with per-launch device time, (2) one --set full capture of the forward kernel
cd $GRAFT_REPO_ROOT
make -C oracle -s
CMD="python bench.py --steps 2 --warmup 1"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/plain_bench.log | cut -c1-400
$CMD > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 1 -c 1 -o gpurun_out/r1_forward_full $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
