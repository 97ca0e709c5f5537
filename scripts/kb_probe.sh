#!/bin/bash
cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_edge.py tests/test_gpu_properties.py -x -q 2>&1 | tail -3
export REPS=7
for v in base kb3 base kb3; do
  export HMC_B200_LIB=$PWD/build/var/libhmc_$v.so
  echo "--- $v"; timeout 300 python scripts/probe_count.py 2000 1000 0.05 0 10 2000 2>&1 | tail -1
done
