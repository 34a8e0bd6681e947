#!/bin/bash
# scaling sweep on one box: python -m torch.distributed.run per N, as the driver does
mkdir -p gpurun_out
for n in ${@:-1 2 4 8}; do
  if [ "$n" = "1" ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/scale_n$n.json") if x.startswith("{")][-1]; d=json.loads(l)
    print("N=%d  %.3f ms/step  %.3f Gsamples/s  e2e %.3f ms  kernel %.3f ms  mode %s" % (d["n_gpus"], d["ms_per_step"], d["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["mode"]))
except Exception as e:
    print("N=$n failed:", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
