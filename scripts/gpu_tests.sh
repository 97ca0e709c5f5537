cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv
make -C oracle -s 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q 2>&1 | tail -30
