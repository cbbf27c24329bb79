#!/bin/bash
# Round-2 evidence run (one gpurun call): bench lines, launch list, ncu --set full captures of the dominant kernels.
# Every ncu capture follows a plain run of the same command that exited 0.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_final.err
python bench.py --workload multiset > gpurun_out/r02_bench_multiset_1gpu.json 2>> gpurun_out/r02_bench_final.err
python tools/bench_configs.py > gpurun_out/r02_bench_configs.jsonl 2>&1
python tools/bench_chains.py > gpurun_out/r02_bench_chains.jsonl 2>&1
python tools/bench_single_chain.py 1 1000 2>/dev/null | head -1 > gpurun_out/r02_bench_single_chain.json
# launch list of the bench command (cold-cache, serialised: compare shares)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --chain-steps 20 --chain-burn-in 5 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --chain-steps 20 --chain-burn-in 5 > /dev/null 2>&1
# full captures
python tools/phase_profile.py --defects 3 --chains 4096 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:eval_warp -s 3 -c 1 -o gpurun_out/r02f_sector -f python tools/phase_profile.py --defects 3 --chains 4096 > /dev/null 2>&1
python tools/run_d32.py 592 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:eval_cta_orbit -s 1 -c 1 -o gpurun_out/r02f_d32_orbit -f python tools/run_d32.py 592 > /dev/null 2>&1
python tools/pack2_check.py --chains 16384 --reps 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:eval_pack2 -s 1 -c 1 -o gpurun_out/r02f_pack2 -f python tools/pack2_check.py --chains 16384 --reps 2 > /dev/null 2>&1
python tools/bench_chains.py 3:4096:20 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:chains_ -s 20 -c 2 -o gpurun_out/r02f_chain_kernels -f python tools/bench_chains.py 3:4096:20 > /dev/null 2>&1
rm -f gpurun_out/pack2_*.npz
# summaries on the box (the reports themselves are too large to travel back)
python tools/ncu_summary.py gpurun_out/r02f_sector.ncu-rep gpurun_out/r02_ncu_full_warp_cheb_parity_sector_realH_d16_final_summary.txt "round 2 final: d = 16 sector kernel, 4096 models (tools/phase_profile.py), ncu --set full --clock-control none" > /dev/null
python tools/ncu_summary.py gpurun_out/r02f_d32_orbit.ncu-rep gpurun_out/r02_ncu_full_cta_orbit_parity_realH_d32_summary.txt "round 2: d = 32 orbit kernel (eml_cta_orbit.cu), 592 models (tools/run_d32.py 592), ncu --set full --clock-control none" > /dev/null
python tools/ncu_summary.py gpurun_out/r02f_pack2.ncu-rep gpurun_out/r02_ncu_full_pack2_realH_d4_summary.txt "round 2: d = 4 two-models-per-warp kernel (eval_pack2_kernel), 16384 models, ncu --set full --clock-control none" > /dev/null
python tools/ncu_summary.py gpurun_out/r02f_chain_kernels.ncu-rep gpurun_out/r02_ncu_full_chain_kernels_summary.txt "round 2: warp-per-chain proposal / accept kernels, 4096 d = 16 chains (one group), ncu --set full --clock-control none" > /dev/null
python tools/ncu_lines.py gpurun_out/r02f_d32_orbit.ncu-rep effective_model_learning_b200/csrc/eml_cta_orbit.o eval_cta_orbit_kernelILb1 40 effective_model_learning_b200/csrc/eml_cta_orbit.cu 2>&1 | cut -c1-220 > gpurun_out/r02_ncu_lines_cta_orbit_d32.txt
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out
