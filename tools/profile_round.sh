#!/bin/bash
# Round profile on one B200 (run through gpurun): bench (all BASELINE configs), reference arm, ncu launch lists, ncu --set full
# of the tcgen05 GEMM launches of one bench step and of the attention / recurrent kernels.  Each ncu pass runs only after the
# plain command exited 0.
R=${1:-r02}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${R}.json 2> gpurun_out/bench_ref_${R}.err
Q="--no-cpu-baseline --no-e2e --no-widened --no-configs"
python bench.py --layers 5 --steps 10 --warmup 3 $Q > gpurun_out/bench_${R}_L5.json 2> gpurun_out/bench_${R}_L5.err
python bench.py --hires --batch 512 --steps 10 --warmup 3 $Q > gpurun_out/bench_${R}_hires.json 2> gpurun_out/bench_${R}_hires.err
python bench.py --batch 256 --steps 20 --warmup 3 $Q > gpurun_out/bench_${R}_B256.json 2> gpurun_out/bench_${R}_B256.err
python bench.py --mode train --batch 128 --steps 20 --warmup 5 --no-configs > gpurun_out/bench_${R}_train128.json 2> gpurun_out/bench_${R}_train128.err
python bench.py --mode train --batch 32 --steps 20 --warmup 5 --no-configs > gpurun_out/bench_${R}_train32.json 2> gpurun_out/bench_${R}_train32.err
python bench.py --mode train --real-head --batch 32 --steps 10 --warmup 3 --no-configs > gpurun_out/bench_${R}_train32_bilstm.json 2> gpurun_out/bench_${R}_train32_bilstm.err
# ---- launch lists (cold-cache, serialised: compare shares) ----
python bench.py --steps 1 --warmup 3 $Q > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 $Q > gpurun_out/ncu_launches_${R}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${R}_train128.csv \
    python bench.py --mode train --batch 128 --steps 1 --warmup 3 --no-graph --no-configs > gpurun_out/ncu_launches_${R}_train.log 2>&1
# ---- ncu --set full: the 15 GEMM launches of the first timed step (3 warm-up steps x 15 launches skipped) ----
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 --launch-skip 45 -c 15 -f \
    -o gpurun_out/gemm_${R} python bench.py --steps 1 --warmup 3 $Q > gpurun_out/ncu_gemm_${R}.log 2>&1
tail -2 gpurun_out/ncu_gemm_${R}.log
export_raw() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null && rm -f gpurun_out/$1.ncu-rep; }   # 64 MiB limit on what travels back
export_raw gemm_${R}
# ---- the HBM-side kernels of the same step: one launch each ----
ncu --set full --clock-control none --import-source on \
    -k regex:'region_rows|cast_f32_bf16|cross_attn_tcgen05|layernorm_kernel|ln_blend|i2t_pool|viterbi16|splitk_reduce_ln' \
    --launch-skip 42 -c 14 -f -o gpurun_out/hbm_${R} python bench.py --steps 1 --warmup 3 $Q > gpurun_out/ncu_hbm_${R}.log 2>&1
export_raw hbm_${R}
# ---- hi-res attention (wide tcgen05 variant, P from TMEM) ----
ncu --set full --clock-control none --import-source on -k regex:cross_attn_tcgen05 --launch-skip 3 -c 1 -f \
    -o gpurun_out/attn_wide2_${R} python bench.py --hires --batch 512 --steps 1 --warmup 3 $Q > gpurun_out/ncu_attn_${R}.log 2>&1
export_raw attn_wide2_${R}
# ---- training: single-query attention backward, fused BiLSTM step kernels ----
ncu --set full --clock-control none --import-source on -k regex:'attn_sq1|lstm_step|emission_head_bwd|colsum_wide' --launch-skip 40 -c 12 -f \
    -o gpurun_out/train_${R} python bench.py --mode train --real-head --batch 32 --steps 1 --warmup 2 --no-graph --no-configs > gpurun_out/ncu_train_${R}.log 2>&1
export_raw train_${R}
# emission head: stage timings
for b in 256 1024 2048; do python tools/lstm_bench.py $b 128 2>&1 | grep "B="; done > gpurun_out/lstm_bench_${R}.log
python tools/lstm_train_bench.py 16 32 64 128 > gpurun_out/lstm_train_bench_${R}.log 2>&1
du -sh gpurun_out; ls -la gpurun_out/*${R}* | head -40
