#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fusion.py tests/test_gpu_region_tail.py -x -q 2>&1 | tail -2
for n in 1 2; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-widened 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']['region_rows']; print(round(d['value']), d['ms_per_step'], 'region_rows', k['ms_per_step'], k['frac_hbm'])"; done
timeout 300 python bench.py --hires --batch 512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-widened 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']['region_rows']; print('hires', round(d['value']), 'region_rows', k['ms_per_step'], k['frac_hbm'])"
