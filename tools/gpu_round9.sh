#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_inflight.py tests/test_gpu_fusion.py -x -q 2>&1 | tail -2
for n in 1 2 2 1; do timeout 300 python bench.py --inflight $n --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-widened 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('inflight $n', round(d['value']), d['ms_per_step'], d['config']['launch'])"; done
timeout 300 python bench.py --inflight 2 --layers 5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-widened 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('L5 inflight 2', round(d['value']), d['ms_per_step'])"
