#!/bin/bash
# Round-2 artefacts on one GPU: full -m gpu test log, the bench line (driver flags), the reference arm, the ncu launch
# list of the bench command, ncu --set full captures of the headline kernels.  Every step under its own timeout.
tag=r2
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout=150 -s 2>&1 | tail -60 > gpurun_out/${tag}_gputests.log; tail -3 gpurun_out/${tag}_gputests.log
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; wc -c gpurun_out/${tag}_bench_n1.json; tail -2 gpurun_out/${tag}_bench_n1.err
timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_n1.err; tail -c 400 gpurun_out/${tag}_bench_reference.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
timeout 300 bash profiles/run_ncu_one.sh fast ${tag}_prof_fast_final > /dev/null 2>&1
timeout 300 bash profiles/run_ncu_one.sh parity ${tag}_prof_parity_final > /dev/null 2>&1
ls -la gpurun_out/${tag}_*
