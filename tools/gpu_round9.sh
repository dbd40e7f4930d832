#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_emission.py tests/test_gpu_tagging_pipeline.py -x -q 2>&1 | tail -4
for d in 0 0 4 5 37; do echo "== ICKA_LSTM_DEBUG=$d"; for b in 1024; do ICKA_LSTM_DEBUG=$d timeout 120 python tools/lstm_bench.py $b 128 2>&1 | grep "B=" | sed 's/.*| recurrent/recurrent/; s/| classifier.*//'; done; done
