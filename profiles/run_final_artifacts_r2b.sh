#!/bin/bash
# Round-2 (third session) artefacts on one GPU: full -m gpu test log, the bench line (driver flags), the ncu launch list of
# the bench command, ncu --set full captures of the headline kernels, compute-sanitizer.  Every step under its own timeout.
tag=r2
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout=150 -s 2>&1 | tail -60 > gpurun_out/${tag}_gputests.log; tail -3 gpurun_out/${tag}_gputests.log
timeout 500 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; wc -c gpurun_out/${tag}_bench_n1.json; tail -2 gpurun_out/${tag}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
timeout 200 bash profiles/run_ncu_one.sh fast ${tag}_prof_fast_final > /dev/null 2>&1
timeout 300 bash profiles/run_ncu_one.sh parity ${tag}_prof_parity_final > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:julia_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_prof_julia_final python profiles/prof_julia.py > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none -k regex:order_place -s 1 -c 1 -f -o gpurun_out/${tag}_prof_order_place python profiles/prof_order.py > /dev/null 2>&1
for tool in memcheck racecheck; do
  RMDF_ORDER_MIN_ITEMS=64 timeout 400 compute-sanitizer --tool $tool python profiles/sanitize.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload|Error|error" | head -5 > gpurun_out/${tag}_sanitizer_$tool.txt
  cat gpurun_out/${tag}_sanitizer_$tool.txt
done
ls -la gpurun_out/${tag}_*
