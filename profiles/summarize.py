"""Turns the raw ncu artefacts gpurun brought back into the small text/CSV summaries kept in profiles/.
Usage: python profiles/summarize.py <launches.csv> <report.ncu-rep> <tag>"""
import collections
import csv
import subprocess
import sys

launch_csv, rep, tag = sys.argv[1:4]
lines = [l for l in open(launch_csv) if not l.startswith("==")]
tot, cnt = collections.OrderedDict(), collections.Counter()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace(", ", ";")      # template arguments: no commas inside a CSV field
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    v = v / 1000.0 if row["Metric Unit"] == "ns" else (v * 1000.0 if row["Metric Unit"] == "ms" else v)
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
S = sum(tot.values())
out = ["kernel,launches,total_us,avg_us,share_pct"]
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    out.append("%s,%d,%.1f,%.2f,%.1f" % (k, cnt[k], v, v / cnt[k], 100 * v / S))
out.append("TOTAL,%d,%.1f,," % (sum(cnt.values()), S))
open("profiles/%s_launch_summary.csv" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
out = ["kernel," + ",".join(want), "unit," + ",".join(units[idx[h]] if h in idx else "" for h in want)]
for r in rows[2:]:
    out.append(r[idx["Kernel Name"]].split("(")[0].replace(", ", ";") + "," + ",".join(r[idx[h]].replace(",", "") if h in idx else "" for h in want))
open("profiles/%s_top_kernels_ncu.csv" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
