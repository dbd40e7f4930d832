#!/bin/bash
for d in 4 5 7; do echo "== ICKA_LSTM_DEBUG=$d"; for b in 1024 256; do ICKA_LSTM_DEBUG=$d timeout 120 python tools/lstm_bench.py $b 128 2>&1 | grep "B=" | sed 's/.*| recurrent/recurrent/; s/| classifier.*//'; done; done
