#!/bin/bash
# One gpurun call: smoke, bench (N=1), ncu launch list, ncu full capture of the march kernel.
set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
python bench.py --steps 5 --warmup 3 --mode fast --no-cpu-baseline > gpurun_out/bench_n1_fast.json 2>> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1_fast.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/prof_march \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_march.log 2>&1
tail -3 gpurun_out/prof_march.log
ls -la gpurun_out
