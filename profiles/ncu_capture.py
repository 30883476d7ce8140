"""profiles/ncu_capture.json from a `ncu --set full` report: per kernel group the LARGEST captured
launch with its measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum), duration and
the counters that say what bounds it.  bench.py reads the file for `roofline.traffic` / `dram_frac`.

usage:  ncu -i gpurun_out/<x>.ncu-rep --page raw --csv > /tmp/raw.csv
        python profiles/ncu_capture.py /tmp/raw.csv "<command the report was taken on>" [out.json]"""
import csv
import json
import os
import re
import sys

GROUPS = [
    ("assoc_nn", r"assoc_(cells|nn)"),
    ("extract_select", r"extract_select"),
    ("extract_normals", r"extract_normals"),
    ("extract_pack", r"extract_pack"),
    ("map_build", r"map_(insert|alloc|scatter|cells|clear)"),
    ("segment", r"segment_scatter"),
    ("lin_chunk", r"moment_"),
    ("lin_finalize", r"eval_\w+<\(bool\)0>|eval_\w+<false>"),
    ("err_finalize", r"eval_\w+<\(bool\)1>|eval_\w+<true>"),
    ("commit", r"commit_"),
]
METRICS = {
    "duration_ns": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "inst": "smsp__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lanes_active": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
              "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main(path, command, out):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        e = {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
             "id": r[col["ID"]]}
        for k, m in METRICS.items():
            if m in col:
                v = num(r[col[m]])
                if v is not None:
                    v *= UNIT_SCALE.get(units[col[m]], 1.0)
                e[k] = v
        launches.append(e)
    result = {}
    for group, pat in GROUPS:
        cand = [e for e in launches if re.search(pat, e["kernel"]) and e.get("duration_ns")]
        if not cand:
            continue
        e = max(cand, key=lambda x: x["duration_ns"])
        dram = (e.get("dram_read") or 0.0) + (e.get("dram_write") or 0.0)
        result[group] = {
            "kernel": e["kernel"][:120], "grid": e["grid"], "block": e["block"], "launch_id": e["id"],
            "duration_us": round(e["duration_ns"] / 1e3, 2), "dram_bytes": int(dram),
            "dram_GBps": round(dram / e["duration_ns"], 2),  # bytes / ns = GB/s
            "warp_instructions": int(e.get("inst") or 0),
            "issue_active_pct": e.get("issue_active_pct"), "warps_active_pct": e.get("warps_active_pct"),
            "lanes_active_per_instruction": e.get("lanes_active"), "dram_throughput_pct": e.get("dram_pct"),
            "l1_hit_pct": e.get("l1_hit_pct"), "l2_hit_pct": e.get("l2_hit_pct"),
            "fp64_pipe_pct": e.get("fp64_pipe_pct"), "registers": e.get("regs"),
            "launches_captured": len(cand), "command": command,
            "note": "ncu --set full --clock-control none, one launch (cold caches, serialised replay): "
                    "compare bytes and counters, not the absolute time",
        }
    with open(out, "w") as f:
        json.dump(result, f, indent=1)
    for g, e in result.items():
        print(f"{g:16s} {e['duration_us']:9.1f} us  dram {e['dram_bytes'] / 1e6:8.2f} MB ({e['dram_GBps']:7.1f} GB/s)  "
              f"inst {e['warp_instructions']:10d}  issue {e['issue_active_pct']}  warps {e['warps_active_pct']}  "
              f"lanes {e['lanes_active_per_instruction']}  grid {e['grid']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "",
         sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_capture.json"))
