#!/bin/bash
# Second profile pass of round 2 (run on the GPU box through gpurun): the -m gpu suite, the full bench line, a launch list of one
# bench run, and full captures of the kernels added since profiles/r2_* (BGZF decode, zero-copy column staging).  Numbers
# printed under ncu are never bench values.
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_tests.log
timeout 900 python bench.py > gpurun_out/${tag}_final_bench.json 2> gpurun_out/${tag}_final_bench.err; echo "bench rc=$?"
GTB_BENCH_THREADS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
BB_REGIONS=2 BB_THREADS=1 timeout 300 ncu --set full --import-source on --clock-control none \
    -k regex:'bgzf_inflate_kernel|bam_walk_blocks_kernel|bam_walk_kernel|bam_tie_kernel|bam_gather_kernel|bam_classify_kernel' \
    -s 12 -c 6 -o gpurun_out/${tag}_bgzf -f python tools/bgzf_bench.py > gpurun_out/${tag}_bgzf.log 2>&1
GTB_ZERO_COPY=1 GTB_BENCH_THREADS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:'gather_columns_kernel' \
    -s 4 -c 1 -o gpurun_out/${tag}_gather -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_gather.log 2>&1
BB_REGIONS=6 GTB_TRACE=1 timeout 200 python tools/bgzf_bench.py > gpurun_out/${tag}_bgzf_trace.log 2>&1
grep -v "gtb trace" gpurun_out/${tag}_bgzf_trace.log | tail -3
tail -c 600 gpurun_out/${tag}_final_bench.json
ls -la gpurun_out/${tag}_*
