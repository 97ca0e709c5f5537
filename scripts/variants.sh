#!/bin/bash
# usage: variants.sh "<probe args>" lib1.so lib2.so ...   (libraries relative to the repo root)
cd $GRAFT_REPO_ROOT
args=$1; shift
for lib in "$@"; do echo "== $lib"; HMC_B200_LIB=$GRAFT_REPO_ROOT/$lib python scripts/probe_c2.py $args 2>&1 | grep "K="; done
