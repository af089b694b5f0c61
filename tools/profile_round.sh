#!/bin/bash
# Round profile (run on the GPU box through gpurun): launch list of one bench run + one full capture of each hot kernel,
# and one capture of the discovery re-alignment kernel.  Numbers printed by bench.py under ncu are never bench values.
tag=${1:-r2}
mkdir -p gpurun_out
GTB_BENCH_THREADS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
GTB_BENCH_THREADS=1 GTB_CHUNKS=1 ncu --set full --import-source on --clock-control none \
    -k regex:'probe_kernel|chain_kernel|chain_general_kernel|slow_kernel|score_kernel|score_deferred_kernel' \
    -s 18 -c 6 -o gpurun_out/${tag}_hot -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_hot.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'sw_kernel' -s 2 -c 1 -o gpurun_out/${tag}_sw -f \
    python tools/sw_bench.py --pairs 100000 --reps 1 --cpu-sample 0 --min-query 140 > gpurun_out/${tag}_sw.log 2>&1
for f in launches_bench hot sw; do tail -n 1 gpurun_out/${tag}_$f.log | cut -c1-300; done
