#!/bin/bash
# usage: profile_fwd.sh <tag> <N> <L> <missing> <threads> <K>
cd $GRAFT_REPO_ROOT
make -C oracle -s
tag=$1; shift
python scripts/probe_one.py "$@" > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 1 -c 1 -o gpurun_out/prof_$tag python scripts/probe_one.py "$@" > gpurun_out/ncu_$tag.log 2>&1
cat gpurun_out/plain_$tag.log; tail -3 gpurun_out/ncu_$tag.log
