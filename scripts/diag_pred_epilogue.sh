#!/bin/bash
# Ablation of the predictive GEMM's output path in the burst regime (needs the diagnostic build:
#   python -m bayesvlm_b200.build --diag   ->  bayesvlm_b200/libbvlm_diag.so, loaded through BVLM_LIB):
#   BVLM_DEBUG_EPI = 0 full kernel | 1 epilogue math only | 2 + shared-memory staging, no stores | 3 bulk stores of stale slabs | 4 direct stores | 5 as 3 in half-height boxes
#   BVLM_DEBUG_SHORTK = 1: one K block per tile (the kernel is its epilogue)
# gpurun --timeout 900 -- bash scripts/diag_pred_epilogue.sh
export BVLM_LIB=$PWD/bayesvlm_b200/libbvlm_diag.so
for m in 0 1 2 3; do echo "epi_mode=$m"; BVLM_DEBUG_EPI=$m python scripts/pred_burst_time.py 20 3 2>&1 | tail -1; done
for m in 0 1 2 3; do echo "shortk epi_mode=$m"; BVLM_DEBUG_SHORTK=1 BVLM_DEBUG_EPI=$m python scripts/pred_burst_time.py 20 3 2>&1 | tail -1; done
