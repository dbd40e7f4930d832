#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run gemm_debug python tools/gemm_debug.py
run kernels python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm_bf16.py -q -m gpu
run fusion python -m pytest tests/test_gpu_fusion.py -q -m gpu
TAILN=40 run microbench python tools/microbench.py 1024
