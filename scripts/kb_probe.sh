#!/bin/bash
cd $GRAFT_REPO_ROOT
export REPS=7
for v in "" e2 "" e2; do
  if [ -n "$v" ]; then export HMC_B200_LIB=$PWD/build/var/libhmc_$v.so; else unset HMC_B200_LIB; fi
  echo "--- ${v:-pfl1}"; timeout 300 python scripts/probe_count.py 2000 1000 0.05 0 10 2000 2>&1 | tail -1
done
