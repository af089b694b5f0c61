#!/bin/bash
# bench.py (no CPU baseline) under several environment settings: tools/bench_env.sh "GTB_CHAIN_SORT=0" "GTB_CHAIN_SORT=1" ...
for e in "$@"; do
  echo "== $e"
  env $e python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:round(v,4) for k,v in d['kernels_ms'].items()}, 'value %.1fM (%.3f ms) e2e %.1fM (%.3f ms)'%(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']))"
done
