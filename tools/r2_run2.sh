#!/bin/bash
set -x
for v in prof prof12; do
  sl=16; [ $v = prof12 ] && sl=12
  EML_B200_LIB=build/variants/libeml_$v.so timeout 300 python tools/phase_profile.py --slots-per-sm $sl
done
EML_B200_LIB=build/variants/libeml_prof.so timeout 300 python tools/phase_profile.py --slots-per-sm 8 --flags 1 --chains 4096
EML_B200_LIB=build/variants/libeml_prof.so timeout 300 python tools/phase_profile.py --defects 2 --chains 1024 8192
EML_B200_LIB=build/variants/libeml_prof.so timeout 300 python tools/phase_profile.py --defects 1 --chains 4096
for v in sb12 sb14; do EML_B200_LIB=build/variants/libeml_$v.so timeout 300 python tools/phase_profile.py --chains 65536; done
timeout 300 python tools/phase_profile.py --chains 4096 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_warp_mma -s 4 -c 1 -o gpurun_out/r02_sb_v1 python tools/phase_profile.py --chains 4096
