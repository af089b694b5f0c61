#!/bin/bash
# default build (bulk staging on): parity tests of the SW kernel, then one full ncu capture of sw_kernel
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_sw.py -m gpu -q -x --timeout=90 -p no:cacheprovider > gpurun_out/r2d_sw_default_tests.log 2>&1; echo "sw tests (default = bulk) rc=$?"; tail -1 gpurun_out/r2d_sw_default_tests.log
timeout 150 ncu --set full --import-source on --clock-control none -k regex:'sw_kernel' -s 2 -c 1 -o gpurun_out/r2d_sw -f \
    python tools/sw_bench.py --pairs 100000 --reps 1 --cpu-sample 0 --min-query 140 > gpurun_out/r2d_sw.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2d_sw.ncu-rep
