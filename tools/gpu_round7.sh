#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_emission.py -x -q -s > gpurun_out/pytest_r01e_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r01e_new.log
tail -12 gpurun_out/pytest_r01e_new.log
for b in 1024 256 2048 128; do timeout 120 python tools/lstm_bench.py $b 128 2>&1 | grep "B="; done > gpurun_out/lstm_bench.log
cat gpurun_out/lstm_bench.log
for d in 1 3; do echo "== ICKA_LSTM_DEBUG=$d"; for b in 1024 256; do ICKA_LSTM_DEBUG=$d timeout 120 python tools/lstm_bench.py $b 128 2>&1 | grep "B=" | sed 's/.*| recurrent/recurrent/; s/| classifier.*//'; done; done > gpurun_out/lstm_probe.log 2>&1
cat gpurun_out/lstm_probe.log
