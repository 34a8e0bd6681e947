#!/bin/bash
# usage: run_ncu_one.sh <fast|parity> <outname>
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/$2 python profiles/prof_one.py $1 2 3 2>&1 | tail -2
