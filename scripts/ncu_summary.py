"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).  Usage:
    python scripts/ncu_summary.py <tag> <launches.csv> <prof.ncu-rep>
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
out = Path("profiles") / f"{tag}.md"
lines = [f"# ncu summary `{tag}`", "", "## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)",
         "per-launch times are cold-cache and serialised: compare shares, not absolutes", "", "| ms total | share | launches | kernel |", "|---:|---:|---:|---|"]
rows = [l for l in open(launches) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(rows):
    agg.setdefault(row["Kernel Name"].split("(")[0][:90], []).append(float(row["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
mine = {k: v for k, v in agg.items() if "ngm::" in k}
if not mine:            # launch list taken with `--kernel-name-base demangled -k regex:ngm::` : only this library's kernels, namespace stripped
    mine = {k: v for k, v in agg.items() if "cub::" not in k}
tot_mine = sum(sum(v) for v in mine.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
    lines.append(f"| {sum(v) / 1e6:.3f} | {100 * sum(v) / tot:.1f}% | {len(v)} | `{k}` |")
lines += ["", f"share among this library's kernels only (total {tot_mine / 1e6:.3f} ms):", ""]
for k, v in sorted(mine.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"* `{k}`: {100 * sum(v) / tot_mine:.1f}% ({sum(v) / len(v) / 1e3:.1f} us per launch)")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines += ["", "## `ncu --set full --clock-control none` per kernel", ""]
for row in r[2:]:
    lines.append(f"### `{row[hdr.index('Kernel Name')].split('(')[0]}`")
    lines.append("")
    for w in want:
        if w in hdr:
            lines.append(f"* {w} = {row[hdr.index(w)]} {units[hdr.index(w)]}")
    lines.append("")
out.write_text("\n".join(lines) + "\n")
print(out)
