#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_emission.py tests/test_gpu_tagging_pipeline.py -x -q 2>&1 | tail -3
for b in 128 256 1024 2048; do timeout 120 python tools/lstm_bench.py $b 128 2>&1 | grep "B="; done > gpurun_out/lstm_bench_r01e.log
cat gpurun_out/lstm_bench_r01e.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lstm_rec --launch-skip 2 -c 1 --csv --log-file gpurun_out/ncu_lstm_small.csv python tools/lstm_bench.py 128 128 > gpurun_out/ncu_lstm.log 2>&1
tail -2 gpurun_out/ncu_lstm_small.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ncu_smoke.csv python __graft_entry__.py smoke > gpurun_out/ncu_smoke.log 2>&1; echo "smoke under ncu rc=$?"; tail -3 gpurun_out/ncu_smoke.log
