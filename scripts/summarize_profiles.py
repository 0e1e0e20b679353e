"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (launch shares + key metrics)."""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
launch_csv = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches_r1.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/flow_r1_umma_b512.ncu-rep"
what = sys.argv[4] if len(sys.argv) > 4 else "python bench.py --steps 20 --warmup 3"

rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    val = float(d["Metric Value"].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1)
    agg[d["Kernel Name"]][0] += 1
    agg[d["Kernel Name"]][1] += val * scale
tot = sum(v[1] for v in agg.values())
out = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 {what}`",
       "# per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes", "",
       "| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k[:100]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
open(f"profiles/{tag}_launches.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum ", "sm__cycles_elapsed.avg", "smsp__pcsamp_warps_issue_stalled", "sm__inst_executed_pipe_uniform", "sm__pipe_tensor_subpipe",
        "lts__t_sectors_srcunit_tex_op_read.sum ", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed_pipe_tmem", "sm__inst_executed_pipe_tc"]
lines = [f"# ncu --set full --clock-control none, one launch of the flow kernel ({rep.split('/')[-1]})"]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k.strip()) for k in keys) and "not_issued" not in h:
        lines.append(f"{h} | {u} | {v}")
open(f"profiles/{tag}_flow_umma_b512_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
print(len(lines), "metric lines")
