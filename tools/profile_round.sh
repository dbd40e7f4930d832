#!/bin/bash
# Round profile on one B200 (run through gpurun): bench, reference arm, ncu launch list, ncu --set full of the
# tcgen05 GEMM launches of one bench step.  Each ncu pass runs only after the plain command exited 0.
R=${1:-r01}
mkdir -p gpurun_out
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${R}.json 2> gpurun_out/bench_ref_${R}.err
python bench.py --layers 5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_L5.json 2> gpurun_out/bench_${R}_L5.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-widened > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-widened > gpurun_out/ncu_launches_${R}.log 2>&1
# 3 warm-up steps x 15 GEMM launches are skipped; the 15 launches of the first timed step are captured
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 --launch-skip 45 -c 15 -f \
    -o gpurun_out/gemm_${R} python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-widened > gpurun_out/ncu_gemm_${R}.log 2>&1
tail -2 gpurun_out/ncu_gemm_${R}.log
# emission head: stage timings, then one ncu --set full capture of the persistent recurrent kernel
for b in 256 1024 2048; do python tools/lstm_bench.py $b 128 2>&1 | grep "B="; done > gpurun_out/lstm_bench_${R}.log
# (ncu cannot replay the cooperative launch of 2-CTA clusters of variant 2: the capture is of the single-CTA variant)
ICKA_LSTM_VARIANT=1 ncu --set full --clock-control none --import-source on -k regex:lstm_rec --launch-skip 2 -c 1 -f \
    -o gpurun_out/lstm_${R} python tools/lstm_bench.py 1024 128 > gpurun_out/ncu_lstm_${R}.log 2>&1
python bench.py --hires --batch 512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_hires.json 2> gpurun_out/bench_${R}_hires.err
