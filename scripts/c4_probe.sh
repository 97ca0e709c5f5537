#!/bin/bash
cd $GRAFT_REPO_ROOT
export HMC_SKIP_ORACLE=1
echo "--- early<=8192"; timeout 300 python scripts/named_shapes.py C4 2>&1 | grep -E "kernel_ms|Error" | sed 's/.*"kernel_ms"/kernel_ms/' | cut -c1-200
REPS=5 timeout 300 python scripts/probe_count.py 2000 1000 0.05 0 10 2000 2>&1 | tail -1 | sed 's/{.*//'
