#!/bin/bash
mkdir -p gpurun_out
ICKA_LSTM_DEBUG=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_rec --launch-skip 2 -c 1 -f \
    -o gpurun_out/lstm_r01e_dbg1 python tools/lstm_bench.py 1024 128 > gpurun_out/ncu_lstm.log 2>&1
tail -2 gpurun_out/ncu_lstm.log
