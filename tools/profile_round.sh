#!/bin/bash
# Round profile (run on the GPU box through gpurun): launch list of one bench run + one full capture of each hot kernel.
# Numbers printed by bench.py under ncu are never bench values.
tag=${1:-r1c}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
GTB_CHUNKS=1 ncu --set full --import-source on --clock-control none \
    -k regex:'probe_kernel|chain_kernel|slow_kernel|huge_kernel|score_kernel|prep_flags_kernel|prep_fill_kernel' \
    -s 14 -c 7 -o gpurun_out/${tag}_hot -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_hot.log 2>&1
for f in launches_bench hot; do tail -n 1 gpurun_out/${tag}_$f.log | cut -c1-300; done
