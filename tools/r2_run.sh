#!/bin/bash
set -x
timeout 300 python tools/accuracy_sweep.py || exit 1
timeout 600 python -m pytest tests/test_gpu_cheb_edges.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -4
timeout 300 python tools/bench_configs.py 2>&1 | grep -E "d=4|d=8"
for v in main par12; do
  if [ $v = main ]; then unset EML_B200_LIB; else export EML_B200_LIB=build/variants/libeml_$v.so; fi
  timeout 300 python tools/phase_profile.py --flags 1 --chains 4096 65536 --slots-per-sm 8 | cut -c1-200
done
EML_B200_LIB=build/variants/libeml_par12.so timeout 600 python -m pytest tests/test_gpu_cheb_edges.py -m gpu -q 2>&1 | tail -3
