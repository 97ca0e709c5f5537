#!/bin/bash
# usage: sweep.sh lib.so "<probe args>" cfg...   -- kernel configurations of a -DFWD_SWEEP build
cd $GRAFT_REPO_ROOT
lib=$1; args=$2; shift; shift
for c in "$@"; do echo "== cfg $c"; HMC_B200_FWD_CFG=$c HMC_B200_LIB=$GRAFT_REPO_ROOT/$lib python scripts/probe_c2.py $args 2>&1 | grep "K="; done
