#!/bin/bash
# round-2 measurement batch (one gpurun call)
set -x
timeout 600 python -m pytest tests/test_gpu_chain_statistics.py tests/test_gpu_chains.py -m gpu -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err; tail -c 1500 gpurun_out/r02_bench_v4.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_v4.json 2> gpurun_out/r02_bench_reference_v4.err; tail -c 800 gpurun_out/r02_bench_reference_v4.err
timeout 600 python bench.py --workload multiset --steps 200 --warmup 5 > gpurun_out/r02_bench_multiset_n1.json 2> gpurun_out/r02_bench_multiset_n1.err; tail -c 1500 gpurun_out/r02_bench_multiset_n1.err
timeout 600 python tools/chain_statistics.py 192
timeout 300 python tools/phase_profile.py --chains 4096 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_warp_mma -s 4 -c 1 -o gpurun_out/r02_sb_v2 python tools/phase_profile.py --chains 4096 > /dev/null 2>&1
