#!/bin/bash
for b in 1024 2048 1024 2048; do timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-widened 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('B=$b', round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
