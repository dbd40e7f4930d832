#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 5 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 6 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_v0.json 2> gpurun_out/bench_err.log; echo "bench exit $?"; cat gpurun_out/bench_r01_v0.json; tail -n 5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01_v0.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_ref_r01_v0.json
