#!/bin/bash
# Last check of the round (short): the bulk-copy staging of sw_kernel against the parity sets, its speed next to the cp.async
# staging, and the drop-in / BGZF tests on the final shim.
mkdir -p gpurun_out
GTB_SW_BULK=1 timeout 150 python -m pytest tests/test_gpu_sw.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/r2d_sw_bulk_tests.log 2>&1; echo "bulk tests rc=$?"; tail -2 gpurun_out/r2d_sw_bulk_tests.log
for b in 0 1 0 1; do GTB_SW_BULK=$b timeout 80 python tools/sw_bench.py --pairs 100000 --reps 5 --cpu-sample 0 --min-query 140 2>&1 | tail -1 | cut -c1-260 | sed "s/^/bulk=$b /"; done | tee gpurun_out/r2d_sw_bulk_bench.log
timeout 200 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_bgzf.py -m gpu -q -x --timeout=180 -p no:cacheprovider > gpurun_out/r2d_final_tests.log 2>&1; echo "dropin+bgzf rc=$?"; tail -2 gpurun_out/r2d_final_tests.log
