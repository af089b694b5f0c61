#!/bin/bash
# Runs bench.py (no CPU baseline) for the default library and every tuning build given; prints the kernel times.
for v in "" "$@"; do
  if [ -z "$v" ]; then lib=graphtyper_b200/libgtb200.so; else lib=graphtyper_b200/libgtb200_$v.so; fi
  echo "== ${v:-default}"
  GTB_LIB=$lib python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:round(v,4) for k,v in d['kernels_ms'].items()}, 'value %.1fM e2e %.1fM (%.3f ms)'%(d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['ms_per_step']))"
done
