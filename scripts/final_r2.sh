#!/bin/bash
# last evidence pass of the round: bench line, ncu captures of the final kernel sources (the launch list / GPU suite: see git history of this script)
cd $GRAFT_REPO_ROOT
make -C oracle -s
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench_N1.json 2> gpurun_out/r2_final_bench.err || exit 1
CMD="python bench.py --steps 2 --warmup 1"
export HMC_BENCH_EPASS_REPS=1 HMC_BENCH_SKIP_EPASS=1
ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 2 -c 1 -o gpurun_out/r2_final_forward_full $CMD > gpurun_out/ncu_f2.log 2>&1
tail -1 gpurun_out/ncu_f2.log
ncu --set full --clock-control none --import-source on -k regex:kbest_kernel -s 2 -c 1 -o gpurun_out/r2_final_kbest_full $CMD > gpurun_out/ncu_f3.log 2>&1
tail -1 gpurun_out/ncu_f3.log
