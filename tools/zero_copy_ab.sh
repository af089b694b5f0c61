#!/bin/bash
# one GPU call: full -m gpu suite, then the bench (no CPU arm) with the staged and the zero-copy column paths
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/zc_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/zc_tests.log
for z in 0 1; do
  GTB_ZERO_COPY=$z timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/zc_bench_z$z.json 2> gpurun_out/zc_bench_z$z.err
  python - <<P
import json
d=json.loads(open("gpurun_out/zc_bench_z$z.json").read().strip().splitlines()[-1])
print("zero_copy=$z value %.1fM (%.3f ms) e2e %.1fM (%.3f ms) launches %s alone %s"%(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['gpu_launches'], d['latency']))
P
done
