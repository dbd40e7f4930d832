#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -8 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r01e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01e.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d['widened'], indent=1))"
