python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_under_ncu.log 2>&1
bash profiles/run_ncu_one.sh fast r1_prof_fast_final > /dev/null 2>&1
python profiles/measure_all.py r1 > gpurun_out/r1_measure_all.log 2>&1; cp profiles/r1_measurements.json gpurun_out/
python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
tail -c 300 gpurun_out/r1_bench_n1.json
