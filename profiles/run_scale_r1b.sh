set -x
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r1_scale_mb_n$n.json 2> gpurun_out/r1_scale_mb_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29718 bench.py --gpus 8 --steps 20 --warmup 3 --workload julia8k > gpurun_out/r1_scale_julia_n8.json 2> gpurun_out/r1_scale_julia_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 4 --steps 20 --warmup 3 --workload julia8k > gpurun_out/r1_scale_julia_n4.json 2> gpurun_out/r1_scale_julia_n4.err
python - <<'PY'
import json
for f in ("mb_n2","mb_n4","mb_n8","julia_n4","julia_n8"):
    try:
        d=json.load(open("gpurun_out/r1_scale_%s.json"%f)); print(f, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), round(d["roofline"]["kernel_ms"],3), d["config"]["parallelism"][:60])
    except Exception as e: print(f, "ERR", e)
PY
