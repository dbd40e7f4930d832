#!/bin/bash
# Developer probe of the cluster LayerNorm GEMM (csrc/gemm_ln_sm100.cu): which part of the epilogue is exposed?
# ICKA_LN_DEBUG bits: 1 = no residual loads, 2 = no stores, 4 = no statistics exchange, 8 = no pass 2 (results are wrong).
for d in ${LN_PROBE_FLAGS:-0 1 2 4 8 7}; do
  echo "== ICKA_LN_DEBUG=$d"
  ICKA_LN_DEBUG=$d timeout 200 python tools/ln_gemm_bench.py 2>&1 | grep -E "^M=|clusters of" | sed 's/ | single_cta[^|]*//'
done
