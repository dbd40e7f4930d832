#!/bin/bash
timeout 500 python -m pytest tests/test_gpu_emission.py tests/test_gpu_tagging_pipeline.py -x -q 2>&1 | tail -3
for b in 1024 1536 2048 4096; do timeout 120 python tools/lstm_bench.py $b 128 2>&1 | grep "B=" | sed 's/: cast.*| recurrent/ recurrent/; s/| classifier.*| module/| module/'; done
