#!/bin/bash
cd $GRAFT_REPO_ROOT
make -C oracle -s
export HMC_BENCH_SKIP_EPASS=1
CMD="python bench.py --steps 2 --warmup 1"
$CMD > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 1 -c 1 -o gpurun_out/r1b_forward_full $CMD > gpurun_out/ncu_b.log 2>&1
tail -1 gpurun_out/ncu_b.log
CMD2="python scripts/probe_estep.py 2000 300 0.05 300"
$CMD2 > gpurun_out/plain_e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:estep_kernel -c 1 -o gpurun_out/r1b_estep_full $CMD2 > gpurun_out/ncu_e.log 2>&1
tail -1 gpurun_out/plain_e.log; tail -1 gpurun_out/ncu_e.log
