for v in ${VARIANTS:-""}; do
  if [ "$v" = main ]; then unset EML_B200_LIB; name=main; else export EML_B200_LIB=build/variants/libeml_$v.so; name=$v; fi
  timeout 100 python bench.py --no-cpu-baseline --no-chains --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']), round(d['roofline']['kernel_ms'],4))"
  timeout 100 python tools/accuracy_sweep.py 2>&1 | grep "d=16"
done
