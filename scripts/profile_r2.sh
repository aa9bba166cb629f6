#!/bin/bash
# Round-2 ncu evidence, one gpurun call:  gpurun --timeout 1500 -- bash scripts/profile_r2.sh
# Launch list of the default bench command (cold-cache, serialised: compare SHARES) + one `--set full` capture per hot kernel.
set -u
O=gpurun_out
mkdir -p $O
if [ "${1:-all}" = "all" ]; then
python bench.py --quick --steps 20 --warmup 3 > $O/r2_bench_quick.json 2> $O/r2_bench_quick.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_ncu_bench.log 2>&1
fi
cap() {  # name, kernel regex, skip, script...
  local name=$1 regex=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s $skip -c 1 -f -o $O/r2_$name "$@" > $O/r2_$name.log 2>&1
}
cap pred_gemm   'gemm2_tn_kernel.*EpiPredictive'   2 python scripts/run_pred_once.py 3
cap pred_quad   'gemm2_tn_kernel.*EpiQuadformPrep' 2 python scripts/run_pred_once.py 3
cap pred_prep   'k_predictive_prep'                2 python scripts/run_pred_once.py 3
cap probit      'k_probit_softmax'                 0 python scripts/run_probit_once.py
cap ggn_rowstats 'gemm2_tn_kernel.*EpiRowLse'      1 python scripts/run_ggn_once.py 2
cap ggn_weights 'gemm2_tn_kernel.*EpiGgnWeights'   1 python scripts/run_ggn_once.py 2
cap ggn_moments 'gemm2_tn_kernel.*EpiStoreF32.*bool.0, .bool.1' 1 python scripts/run_ggn_once.py 2
cap syrk        'gemm2_tn_kernel.*EpiStoreF32.*bool.1, .bool.1' 1 python scripts/run_syrk_once.py 2
cap epig_joint  'gemm2_tn_kernel.*EpiEpigJoint'    0 python scripts/run_epig_once.py 1
cap epig_prepare 'k_epig_prepare'                  0 python scripts/run_epig_once.py 1
ls -la $O/r2_*.ncu-rep
