"""Per-kernel summary of profiles/ncu_metrics.sh output (ncu --csv, one row per metric)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()
for r in rows[1:]:
    if r[col["ID"]] == "ID":
        continue
    k = (r[col["ID"]], r[col["Kernel Name"]].split("(")[0])
    launches.setdefault(k, {})[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
agg = collections.defaultdict(list)
for (_, name), m in launches.items():
    agg[name].append(m)
have_mem = any("dram__bytes_read.sum" in m for ms in agg.values() for m in ms)
print(f"{'kernel':34s} {'n':>4s} {'us/launch':>9s} {'tot_ms':>7s} {'Minst':>7s} {'inst/CTA':>9s} {'issue%':>6s} {'warps%':>6s} {'lanes':>5s}"
      + (f" {'fp64%':>5s} {'L1hit':>5s} {'L2hit':>5s} {'dramMB':>7s}" if have_mem else ""))
for name, ms in sorted(agg.items(), key=lambda kv: -sum(m["gpu__time_duration.sum"] for m in kv[1])):
    n = len(ms)
    t = sum(m["gpu__time_duration.sum"] for m in ms)
    tus = t / 1e3  # ncu reports gpu__time_duration in ns in --csv mode
    inst = sum(m["smsp__inst_executed.sum"] for m in ms)
    blocks = sum(m["launch__grid_size"] for m in ms)
    w = lambda key: sum(m.get(key, 0) * m["gpu__time_duration.sum"] for m in ms) / t
    line = (f"{name[:34]:34s} {n:4d} {tus / n:9.1f} {tus / 1e3:7.2f} {inst / 1e6:7.1f} {inst / blocks:9.0f} "
            f"{w('smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {w('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
            f"{w('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f}")
    if have_mem:
        dram = sum(m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0) for m in ms)
        line += (f" {w('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} {w('l1tex__t_sector_hit_rate.pct'):5.1f} "
                 f"{w('lts__t_sector_hit_rate.pct'):5.1f} {dram / n / 1e6:7.2f}")
    print(line)
