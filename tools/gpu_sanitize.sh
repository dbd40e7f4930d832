#!/bin/bash
# compute-sanitizer memcheck over the small cases of the kernels added this session
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest -x -q tests/test_gpu_ner.py tests/test_gpu_region_tail.py \
  "tests/test_gpu_emission.py::test_bf16_persistent_kernel[128-3-1]" "tests/test_gpu_emission.py::test_bf16_persistent_kernel[128-3-2]" \
  "tests/test_gpu_emission.py::test_bf16_persistent_kernel[1-1-1]" "tests/test_gpu_emission.py::test_bf16_persistent_kernel[1-1-2]" \
  "tests/test_gpu_emission.py::test_bf16_persistent_kernel[640-5-2]" \
  "tests/test_gpu_emission.py::test_emission_head_kernel" "tests/test_gpu_emission.py::test_fp32_per_step_path[3-9-32-32]" \
  > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitize.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|rc=" gpurun_out/sanitize.log | tail -12
