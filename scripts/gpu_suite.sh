#!/bin/bash
# the whole GPU suite + a bench line
cd $GRAFT_REPO_ROOT
make -C oracle -s
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -3 gpurun_out/bench_now.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_now.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['kernel_ms'], d['roofline']['frac'], d['cpu_baseline'])"
