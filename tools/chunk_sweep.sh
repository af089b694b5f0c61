#!/bin/bash
# e2e vs number of pipeline chunks (GTB_CHUNKS) on the bench workload
for c in 1 2 3 4 6 8; do
  GTB_CHUNKS=$c python bench.py --no-cpu-baseline 2>&1 | tail -1 | GTB_C=$c python -c "
import sys, json, os
d = json.loads(sys.stdin.read())
print('chunks', os.environ['GTB_C'], 'value', round(d['value']/1e6,1), 'M/s', 'kernel ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'M/s', round(d['e2e']['ms_per_step'],3), 'ms')"
done
