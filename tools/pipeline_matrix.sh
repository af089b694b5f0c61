#!/bin/bash
# pipeline_probe.py under the queueing knobs (results: profiles/r2b_notes.md): where does the pipelined e2e step lose against the device-resident replay?
export PP_THREADS=4 PP_ONLY=4 PP_STEPS=30 PP_MODES=replay,submit,e2e
for e in "GTB_FIFO=0" "GTB_FIFO=1" "GTB_FIFO=2" "GTB_FIFO=3" "GTB_FIFO=3 PP_THREADS=3 PP_ONLY=3" "GTB_FIFO=3 PP_THREADS=6 PP_ONLY=6" "GTB_FIFO=3 GTB_ZERO_COPY=0" "GTB_FIFO=3 PP_THREADS=2 PP_ONLY=2"; do
  echo "== $e"
  env $e timeout 200 python tools/pipeline_probe.py 2>&1 | grep -v "^\[gtb" | tail -3
done
