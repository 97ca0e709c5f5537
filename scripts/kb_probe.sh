#!/bin/bash
# streamed K-best (beside the forward kernel): register splits forward / K-best
cd $GRAFT_REPO_ROOT
make -C oracle -s
for v in f72_k80 f72_k40 f64_k64; do
  export HMC_B200_LIB=$PWD/build/var/libhmc_$v.so
  echo "--- $v streamed"; timeout 300 python scripts/probe_count.py 2000 1000 0.05 0 10 2000 2>&1 | tail -1
  echo "--- $v serial"; HMC_B200_KBEST_SERIAL=1 timeout 300 python scripts/probe_count.py 2000 1000 0.05 0 10 2000 2>&1 | tail -1
done
unset HMC_B200_LIB
timeout 900 python -m pytest tests/test_gpu_emit.py tests/test_gpu_forward.py tests/test_gpu_edge.py -x -q 2>&1 | tail -15
