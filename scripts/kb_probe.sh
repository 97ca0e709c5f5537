#!/bin/bash
cd $GRAFT_REPO_ROOT
export REPS=7
for v in base dec base dec; do
  export HMC_B200_LIB=$PWD/build/var/libhmc_$v.so
  echo "--- $v"; timeout 300 python scripts/probe_count.py 2000 1000 0.05 0 10 2000 2>&1 | tail -1
done
