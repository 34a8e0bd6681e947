#!/bin/bash
# Scaling sweep on one 8-GPU box, launched per N the way the driver does (torchrun, one rank per GPU).
# usage: bash profiles/run_scale.sh <tag> [N ...]     → gpurun_out/<tag>_scaling.json
tag=${1:-r1}; shift
mkdir -p gpurun_out
python - <<PY > gpurun_out/${tag}_scaling.json
import json, subprocess, sys
res = {}
for wl in ("mandelbulb4k", "julia8k"):
    res[wl] = {}
    for n in [int(x) for x in "${@:-1 2 4 8}".split()]:
        if n == 1:
            cmd = [sys.executable, "bench.py", "--gpus", "1", "--steps", "20", "--warmup", "3", "--no-cpu-baseline", "--workload", wl]
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                   "--master-port", str(29700 + n), "bench.py", "--gpus", str(n), "--steps", "20", "--warmup", "3", "--workload", wl]
        out = subprocess.run(cmd, capture_output=True, text=True)
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not lines:
            res[wl][str(n)] = {"error": out.stderr[-800:]}
            continue
        d = json.loads(lines[-1])
        res[wl][str(n)] = {"ms_per_step": d["ms_per_step"], "value": d["value"], "e2e_ms": d["e2e"]["ms_per_step"],
                           "kernel_ms": d["roofline"]["kernel_ms"], "parallelism": d["config"]["parallelism"], "clocks": d.get("clocks")}
    base = res[wl].get("1", {}).get("ms_per_step")
    for n, v in res[wl].items():
        if base and "ms_per_step" in v:
            v["speedup"] = base / v["ms_per_step"]; v["efficiency"] = v["speedup"] / int(n)
print(json.dumps(res, indent=1))
PY
cat gpurun_out/${tag}_scaling.json | python -c "
import json,sys
d=json.load(sys.stdin)
for wl,r in d.items():
    for n,v in r.items(): print(wl, 'N=%s'%n, {k:(round(x,3) if isinstance(x,float) else x) for k,x in v.items() if k in ('ms_per_step','e2e_ms','kernel_ms','speedup','efficiency','error')})
"
