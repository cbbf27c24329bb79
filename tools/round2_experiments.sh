#!/bin/bash
# The A/B runs DESIGN.md section 8 leaves for the next round, in one gpurun call (about 3 GPU-minutes):
#   tools/build_variant.sh w1 -DEML_WARPS_PER_CTA=1        (build here first: the .so travels with the snapshot)
#   gpurun --timeout 300 -- 'bash tools/round2_experiments.sh > gpurun_out/round2_experiments.log 2>&1'
set -x
python tools/order_pack_ab.py 4096 50                                           # packed model order: identical results + time
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cheb16_cost|pack_models|eval_warp" -c 12 --csv \
    --log-file gpurun_out/launches_pack_v2.csv python tools/order_pack_ab.py 4096 1 > /dev/null 2>&1   # pre-pass kernel times
for g in 2 3 4; do EML_CHAIN_GROUPS=$g python tools/bench_chains.py; done        # chain groups: RJMCMC steps/s at 4096 chains
if [ -f build/variants/libeml_w1.so ]; then                                      # one-warp CTAs: slots released one by one
  for g in 2 4; do EML_B200_LIB=build/variants/libeml_w1.so EML_CHAIN_GROUPS=$g python tools/bench_chains.py; done
  EML_B200_LIB=build/variants/libeml_w1.so python bench.py --no-cpu-baseline --steps 30
fi
