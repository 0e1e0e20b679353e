"""Turns raw ncu outputs (gpurun_out/) into the committed summaries under profiles/.

    python scripts/summarize_profiles.py launches <launches.csv> <out.md> "<profiled command>"
    python scripts/summarize_profiles.py ncu <capture.ncu-rep> <out.txt> ["note"]
    python scripts/summarize_profiles.py stalls <capture.ncu-rep> <out.txt>      # warp-stall samples per source line (top 25)

`ncu` keeps the metrics bench.py and DESIGN.md quote (time, DRAM bytes, L2 -> SM bytes, tensor-pipe activity, occupancy,
shared-memory wavefronts, stall reasons); bench.py reads `dram__bytes_read.sum` / `dram__bytes_write.sum` from these files.
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster",
        "launch__occupancy_limit", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum ", "sm__cycles_elapsed.avg", "smsp__pcsamp_warps_issue_stalled", "sm__inst_executed_pipe_uniform", "sm__pipe_tensor_subpipe",
        "lts__t_sectors_srcunit_tex_op_read.sum ", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed_pipe_tmem", "sm__inst_executed_pipe_tc",
        "smsp__average_warp", "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__cycles_active.avg", "lts__t_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum"]


def launches(path, out, what):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
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
    lines = [f"# ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 {what}`",
             "# per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes", "",
             "| kernel | launches | total us | us / launch | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:110]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


def ncu_summary(rep, out, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# ncu --set full --clock-control none, one launch ({rep.split('/')[-1]}){'; ' + note if note else ''}"]
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k.strip()) for k in KEYS) and "not_issued" not in h:
            lines.append(f"{h} | {u} | {v}")
    open(out, "w").write("\n".join(lines) + "\n")
    print(len(lines), "metric lines ->", out)


def stalls(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    samp = next((h for h in hdr if h.startswith("# Samples") or h == "Warp Stall Sampling (All Samples)"), None)
    src = next((h for h in hdr if h in ("Source", "SASS", "Source Line")), hdr[1])
    if samp is None:
        print("no sampling column in", hdr[:12])
        return
    agg = []
    for r in rows[1:]:
        try:
            agg.append((float(r[col[samp]].replace(",", "") or 0), r[col[src]].strip()[:150], r[0]))
        except (ValueError, IndexError):
            pass
    tot = sum(a[0] for a in agg) or 1.0
    agg.sort(reverse=True)
    lines = [f"# warp-stall samples per source line, top 25 of {int(tot)} ({rep.split('/')[-1]})"]
    lines += [f"{100 * n / tot:5.1f}%  {int(n):7d}  {line}: {text}" for n, text, line in agg[:25]]
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    elif mode == "ncu":
        ncu_summary(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        stalls(sys.argv[2], sys.argv[3])
