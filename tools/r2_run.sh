#!/bin/bash
set -x
timeout 300 python tools/accuracy_sweep.py || exit 1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
EML_B200_LIB=build/variants/libeml_dbg.so timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/bench_single_chain.py
EML_HOST_GRAPH=0 timeout 300 python tools/bench_single_chain.py
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_v5.json 2> gpurun_out/r02_bench_v5.err; tail -c 600 gpurun_out/r02_bench_v5.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v5.json')); r=d['roofline']
print(d['value'], d['e2e']['value'], r['frac'], r['kernel_ms'], r['kernel_ms_after_l2_flush'], d['config']['l2'])"
