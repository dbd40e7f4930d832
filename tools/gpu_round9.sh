#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01h.json 2> gpurun_out/bench_r01h.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01h.json 2> gpurun_out/bench_ref_r01h.err
python bench.py --layers 5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01h_L5.json 2> gpurun_out/bench_r01h_L5.err
python bench.py --hires --batch 512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01h_hires.json 2> gpurun_out/bench_r01h_hires.err
python bench.py --batch 2048 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01h_B2048.json 2> gpurun_out/bench_r01h_B2048.err
python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_r01h_train.json 2> gpurun_out/bench_r01h_train.err
for f in bench_r01h bench_ref_r01h bench_r01h_L5 bench_r01h_hires bench_r01h_B2048 bench_r01h_train; do python -c "
import json,sys; d=json.load(open('gpurun_out/$f.json')); w=d.get('widened') or {}; print('$f', round(d['value']), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), w.get('value'), w.get('stage_ms'))"; done
