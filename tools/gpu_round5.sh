#!/bin/bash
# new rows of this session: NER chunk counts + emission head
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ner.py tests/test_gpu_emission.py -x -q -s > gpurun_out/pytest_r01e_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r01e_new.log
tail -25 gpurun_out/pytest_r01e_new.log
timeout 300 python tools/lstm_bench.py 1024 128 > gpurun_out/lstm_bench.log 2>&1
timeout 300 python tools/lstm_bench.py 256 128 >> gpurun_out/lstm_bench.log 2>&1
timeout 300 python tools/lstm_bench.py 4096 128 >> gpurun_out/lstm_bench.log 2>&1
cat gpurun_out/lstm_bench.log
