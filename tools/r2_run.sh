#!/bin/bash
# round-2 measurement batch (one gpurun call): correctness first, then timing
set -x
EML_B200_DEBUG=1 timeout 300 python tools/accuracy_sweep.py || exit 1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
VARIANTS="${VARIANTS:-main}" bash tools/run_variants.sh
EML_B200_LIB=build/variants/libeml_prof.so timeout 300 python tools/phase_profile.py
EML_B200_LIB=build/variants/libeml_prof.so timeout 300 python tools/phase_profile.py --slots-per-sm 8 --flags 1 --chains 4096 65536
timeout 600 python tools/bench_configs.py
timeout 300 python tools/phase_profile.py --chains 4096 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 12 --csv --log-file gpurun_out/r02_launches_v4.csv python tools/phase_profile.py --chains 4096 > /dev/null 2>&1
