#!/bin/bash
# final round-1 evidence: launch list + full captures of the forward kernel (bench command) and of the E-step kernel
cd $GRAFT_REPO_ROOT
make -C oracle -s
CMD="python bench.py --steps 2 --warmup 1"
$CMD > gpurun_out/plain_f.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv $CMD > gpurun_out/ncu_f1.log 2>&1
export HMC_BENCH_SKIP_EPASS=1
$CMD > gpurun_out/plain_f2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 1 -c 1 -o gpurun_out/r1c_forward_full $CMD > gpurun_out/ncu_f2.log 2>&1
tail -1 gpurun_out/ncu_f2.log
CMD2="python scripts/probe_estep.py 2000 300 0.05 300"
$CMD2 > gpurun_out/plain_e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:estep_kernel -c 1 -o gpurun_out/r1c_estep_full $CMD2 > gpurun_out/ncu_e.log 2>&1
tail -1 gpurun_out/plain_e.log; tail -1 gpurun_out/ncu_e.log
