#!/bin/bash
# Host-side timeline of gtb_submit_reads_multi on the bench workload for several chunk counts (GTB_TRACE)
for c in ${@:-1 3 8}; do
  echo "== chunks $c"
  GTB_TRACE=1 GTB_CHUNKS=$c python tools/e2e_breakdown.py 2>&1 | tail -7
done
