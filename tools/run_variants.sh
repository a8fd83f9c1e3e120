#!/bin/bash
# usage: tools/run_variants.sh v1 v2 ...  (bench each libkmsc_<v>.so; "default" = libkmsc.so)
for v in "$@"; do
  if [ "$v" = default ]; then unset KMSC_LIB; else export KMSC_LIB=$PWD/kmer-sets-compression_b200/libkmsc_$v.so; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3), 'frac', round(d['roofline']['frac'],4), 'W01', d['check']['W01'])"
done
