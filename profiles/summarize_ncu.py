"""Turns an .ncu-rep (ncu --set full --import-source on) into a small tracked summary:
key raw metrics, stall reasons, opcode mix and the hot SASS regions with their lane efficiency.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_xxx.md "title"
"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])


def page(name):
    r = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second"]
lines = ["# %s" % title, "", "Source: `%s` (ncu --set full --clock-control none --import-source on; one launch)." % rep.split("/")[-1], "",
         "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in m:
        lines.append("| `%s` | %s | %s |" % (k, m[k][0], m[k][1]))
lines += ["", "Stall reasons (warps per issue-active cycle, > 0.03):", "", "| reason | value |", "|---|---|"]
for h, (v, u) in sorted(m.items()):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            if float(v) > 0.03:
                lines.append("| %s | %.3f |" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(v)))
        except ValueError:
            pass
src = page("source")
sh, data = src[1], src[2:]
ix = {h: i for i, h in enumerate(sh)}
f = lambda r, k: float(r[ix[k]] or 0)
tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_t = sum(f(r, "Thread Instructions Executed") for r in data)
tot_s = sum(f(r, "# Samples") for r in data) or 1
lines += ["", "SASS: %d instructions, %.3e warp-instructions executed, average active lanes %.2f / 32." % (len(data), tot_i, tot_t / tot_i), ""]
ops = Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    ops[op.split(".")[0]] += f(r, "Instructions Executed")
lines.append("Opcode mix (share of executed warp-instructions): " + ", ".join("%s %.1f%%" % (k, 100 * v / tot_i) for k, v in ops.most_common(14)))
lines += ["", "Hot regions (64-instruction windows with > 1.5 % of executed instructions):", "",
          "| SASS index | share of instructions | active lanes / 32 | share of stall samples |", "|---|---|---|---|"]
for i in range(0, len(data), 64):
    ch = data[i:i + 64]
    ins = sum(f(r, "Instructions Executed") for r in ch)
    if ins / tot_i > 0.015:
        thr = sum(f(r, "Thread Instructions Executed") for r in ch)
        lines.append("| %d–%d | %.1f %% | %.3f | %.1f %% |" % (i, i + len(ch) - 1, 100 * ins / tot_i, thr / ins / 32, 100 * sum(f(r, "# Samples") for r in ch) / tot_s))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
