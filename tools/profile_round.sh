#!/bin/bash
# Round profile (run on the GPU box through gpurun): launch list of one bench run + one full capture of each hot kernel.
# Numbers printed by bench.py under ncu are never bench values.
tag=${1:-r1b}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'probe_kernel|chain_kernel|slow_kernel|huge_kernel|score_kernel' \
    -s 10 -c 5 -o gpurun_out/${tag}_hot -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_hot.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'idx_enum_kernel|idx_group_kernel|idx_table_kernel|idx_heads_kernel' \
    -c 5 -o gpurun_out/${tag}_idx -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_idx.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:sw_kernel -s 1 -c 1 -o gpurun_out/${tag}_sw -f \
    python tools/sw_bench.py --pairs 100000 --reps 1 --cpu-sample 0 --min-query 140 > gpurun_out/${tag}_sw.log 2>&1
for f in hot idx sw; do tail -n 2 gpurun_out/${tag}_$f.log; done
