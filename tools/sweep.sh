#!/bin/bash
# BASELINE configs[2]: inference sweep over the batch per GPU (device-resident `value` only; the default bench line
# carries e2e / widened / cpu_baseline at B=1024).  Usage (on a GPU box): tools/sweep.sh <tag>
tag=${1:-sweep}
mkdir -p gpurun_out
for b in 256 512 2048 4096; do
  timeout 100 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline --no-widened --no-e2e \
    > gpurun_out/bench_${tag}_B${b}.json 2> gpurun_out/bench_${tag}_B${b}.err
  echo "B=$b rc=$?"
done
python - <<PY
import json
for b in (256, 512, 2048, 4096):
    try:
        d = json.load(open(f'gpurun_out/bench_${tag}_B{b}.json'))
        print(b, round(d['value']), round(d['ms_per_step'], 3), round(d['roofline']['frac'], 3), d['clocks'])
    except Exception as e:
        print(b, 'failed', e)
PY
