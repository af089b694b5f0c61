#!/bin/bash
# resident-blocks / store-policy sweep of the re-alignment kernel (GTB_SW_BLOCKS, GTB_SW_STREAM)
for st in 0 1; do for b in ${2:-3 4 6 8}; do
  echo -n "stream=$st blocks=$b: "
  GTB_SW_STREAM=$st GTB_SW_BLOCKS=$b python tools/sw_bench.py --cpu-sample 0 --min-query ${1:-20} 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], d['value'], d['gcups'])"
done; done
