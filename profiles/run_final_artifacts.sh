#!/bin/bash
# Round artefacts: launch list of the bench command, ncu captures of the headline kernels, measurements.
tag=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
bash profiles/run_ncu_one.sh fast ${tag}_prof_fast_final > /dev/null 2>&1
bash profiles/run_ncu_one.sh parity ${tag}_prof_parity_final > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:julia -s 2 -c 1 -f -o gpurun_out/${tag}_prof_julia_final python profiles/prof_julia.py > /dev/null 2>&1
python profiles/measure_all.py ${tag} > gpurun_out/${tag}_measure_all.log 2>&1; cp profiles/${tag}_measurements.json gpurun_out/
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --workload julia8k --steps 20 --warmup 3 > gpurun_out/${tag}_bench_julia_n1.json 2>> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_n1.err
tail -c 600 gpurun_out/${tag}_bench_n1.json; echo; tail -c 300 gpurun_out/${tag}_bench_reference.json
