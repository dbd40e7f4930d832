#!/bin/bash
# First GPU pass: each stage in its own process so one faulting kernel cannot poison the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run crf python -m pytest tests/test_gpu_crf.py -q -m gpu
run kernels python -m pytest tests/test_gpu_kernels.py -q -m gpu
run gemm_debug python tools/gemm_debug.py
run gemm_bf16 python -m pytest tests/test_gpu_gemm_bf16.py -q -m gpu
run fusion python -m pytest tests/test_gpu_fusion.py -q -m gpu
run microbench python tools/microbench.py 1024
