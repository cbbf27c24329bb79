#!/bin/bash
# Final round-2 evidence run (one gpurun call, one GPU): bench lines, launch list, ncu --set full captures of the headline
# evaluator and of the prepare kernel.  Every ncu capture follows a plain run of the same command that exited 0.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_final3.json 2> gpurun_out/r02_bench_final3.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_final3.err
python bench.py --workload multiset > gpurun_out/r02_bench_multiset_1gpu_final3.json 2>> gpurun_out/r02_bench_final3.err
python tools/bench_configs.py > gpurun_out/r02_bench_configs.jsonl 2>&1
python tools/bench_chains.py > gpurun_out/r02_bench_chains.jsonl 2>&1
python tools/bench_single_chain.py 1 1000 2>/dev/null | head -1 > gpurun_out/r02_bench_single_chain.json
python tools/bench_single_chain.py 3 1000 2>/dev/null | head -1 >> gpurun_out/r02_bench_single_chain.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --chain-steps 20 --chain-burn-in 5 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --chain-steps 20 --chain-burn-in 5 > /dev/null 2>&1
python tools/phase_profile.py --defects 3 --chains 4096 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:eval_warp -s 3 -c 1 -o gpurun_out/r02g_sector -f python tools/phase_profile.py --defects 3 --chains 4096 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:prepare_models -s 3 -c 1 -o gpurun_out/r02g_prepare -f python tools/phase_profile.py --defects 3 --chains 4096 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02g_sector.ncu-rep gpurun_out/r02_ncu_full_warp_cheb_parity_sector_realH_d16_final3_summary.txt "round 2 final build (cost-bin lists, self-cleaning fetch workspace, PDL): d = 16 sector kernel, 4096 models (tools/phase_profile.py), ncu --set full --clock-control none" > /dev/null
python tools/ncu_summary.py gpurun_out/r02g_prepare.ncu-rep gpurun_out/r02_ncu_full_prepare_models_d16_final3_summary.txt "round 2 final build: prepare kernel (d = 16, real H, parity; fills the cost-bin lists), 4096 models, ncu --set full --clock-control none (cold caches)" > /dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -15
