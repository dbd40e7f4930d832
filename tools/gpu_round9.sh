#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_emission.py tests/test_gpu_tagging_pipeline.py -x -q 2>&1 | tail -2
timeout 200 python tools/lstm_trace.py 1024 | tail -3
for d in 0 0 5; do echo "== ICKA_LSTM_DEBUG=$d"; ICKA_LSTM_DEBUG=$d timeout 120 python tools/lstm_bench.py 1024 128 2>&1 | grep "B=" | sed 's/.*| recurrent/recurrent/; s/| classifier.*//'; done
