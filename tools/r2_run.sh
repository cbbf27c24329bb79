#!/bin/bash
# round-2 measurement batch (one gpurun call): correctness first, then A/B of the built variants
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
EML_B200_DEBUG=1 timeout 300 python tools/accuracy_sweep.py
VARIANTS="${VARIANTS:-main}" bash tools/run_variants.sh
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/bench_configs.py
