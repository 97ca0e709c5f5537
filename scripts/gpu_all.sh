#!/bin/bash
cd $GRAFT_REPO_ROOT
make -C oracle -s
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
