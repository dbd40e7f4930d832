#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=30 run gemm_debug_pair python tools/gemm_debug.py 2
run gemm_tests python -m pytest tests/test_gpu_gemm_bf16.py -q -m gpu -x
TAILN=16 run microbench_single python tools/microbench.py 1024 1
TAILN=16 run microbench_pair python tools/microbench.py 1024 2
