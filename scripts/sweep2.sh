#!/bin/bash
# usage: sweep2.sh "<probe args>" "cfg cfg ..." lib.so ...   -- kernel configurations x libraries of -DFWD_SWEEP builds
cd $GRAFT_REPO_ROOT
args=$1; cfgs=$2; shift; shift
for lib in "$@"; do for c in $cfgs; do echo -n "$lib cfg $c: "; HMC_B200_FWD_CFG=$c HMC_B200_LIB=$GRAFT_REPO_ROOT/$lib python scripts/probe_c2.py $args 2>&1 | grep "K=" | cut -c1-60; done; done
