#!/bin/bash
# Developer probes of the persistent LSTM kernel (ICKA_LSTM_DEBUG bits; results are WRONG by construction, timings only):
#   1 no dependency waits, 2 no cell arithmetic / stores, 4 relaxed publish, 16 no MMAs, 32 no state stores
R=${1:-r01}
mkdir -p gpurun_out
for d in 0 1 3 4 5 7 37; do
  echo "== ICKA_LSTM_DEBUG=$d"
  for b in 1024 256; do
    ICKA_LSTM_DEBUG=$d timeout 120 python tools/lstm_bench.py $b 128 2>&1 | grep "B=" | sed 's/: cast.*| recurrent/ recurrent/; s/| classifier.*//'
  done
done > gpurun_out/lstm_probe_${R}.log 2>&1
for b in 128 256 1024 2048 4096; do timeout 120 python tools/lstm_bench.py $b 128 2>&1 | grep "B="; done > gpurun_out/lstm_bench_${R}.log
cat gpurun_out/lstm_probe_${R}.log gpurun_out/lstm_bench_${R}.log
