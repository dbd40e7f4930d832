#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 5 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; echo "bench exit $?"; cat gpurun_out/bench_cur.json; tail -n 3 gpurun_out/bench_cur.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tail -n 3 gpurun_out/ncu_launches.log
