#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/prof_march_fast \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --mode fast > gpurun_out/prof_march_fast.log 2>&1
tail -2 gpurun_out/prof_march_fast.log
python bench.py --steps 5 --warmup 3 --mode fast --no-cpu-baseline
