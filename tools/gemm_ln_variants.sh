#!/bin/bash
# A/B timing of the fused GEMM + LayerNorm kernel with parts of its epilogue switched off (MMER_DEBUG_LN_VARIANT bits):
# 1 no residual tile, 2 no z store, 4 no pass 2, 8 no epilogue math, 16 no L2 prefetch of the A operand.  Results of those runs are wrong by construction.
for v in 0 1 4 5 7 15; do
  echo "== variant $v"
  MMER_DEBUG="12=$v" python tools/kernel_bench.py gemm_ln 2>&1 | grep "gemm_ln_fwd"
done
