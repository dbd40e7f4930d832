#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_emission.py -x -q -s > gpurun_out/pytest_r01e_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r01e_new.log
tail -12 gpurun_out/pytest_r01e_new.log
for b in 1024 256 4096 128; do timeout 300 python tools/lstm_bench.py $b 128; done > gpurun_out/lstm_bench.log 2>&1
cat gpurun_out/lstm_bench.log
