#!/bin/bash
set -x
EML_B200_LIB=build/variants/libeml_prof.so timeout 300 python tools/phase_profile.py --chains 148 592 1184 2368
EML_B200_LIB=build/variants/libeml_prof.so timeout 300 python tools/phase_profile.py --chains 148 592 --slots-per-sm 8 --flags 1
EML_B200_LIB=build/variants/libeml_prof.so timeout 300 python tools/phase_profile.py --chains 148 592 --defects 2
