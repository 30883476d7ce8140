"""Per-source-line totals (warp instructions executed, stall samples) of one kernel from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv -k regex:<kernel>` (the report must
have been taken with --import-source on and the kernels compiled with -lineinfo).
usage: python profiles/line_profile.py <csv> [top]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "Line No" and len(r) > 6:
            col = {h: k for k, h in enumerate(r)}
            i_inst, i_samp = col["Instructions Executed"], col["Warp Stall Sampling (All Samples)"]
            i_thr = col["Thread Instructions Executed"]
            per_line, cur, src = {}, None, {}
            j = i + 1
            while j < len(rows) and not (rows[j] and rows[j][0] in ("Line No", "File Name", "Kernel Name")):
                x = rows[j]
                if x and x[0]:
                    if not x[0].isdigit():
                        break
                    cur = int(x[0])
                    src[cur] = x[1]
                elif cur is not None and len(x) > i_thr and x[i_inst] not in ("-", ""):
                    e = per_line.setdefault(cur, [0, 0, 0])
                    e[0] += int(x[i_inst])
                    e[1] += int(x[i_samp] or 0)
                    e[2] += int(x[i_thr])
                j += 1
            tot_i = sum(e[0] for e in per_line.values()) or 1
            tot_s = sum(e[1] for e in per_line.values()) or 1
            if tot_i > 1:
                print(f"== section at csv row {i}: {tot_i} warp instructions, {tot_s} samples, "
                      f"lanes/instr {sum(e[2] for e in per_line.values()) / tot_i:.1f}")
                for ln, e in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
                    print(f"  L{ln:5d} inst {e[0]:9d} {e[0] / tot_i:6.1%}  samples {e[1]:5d} {e[1] / tot_s:6.1%}  "
                          f"lanes {e[2] / max(e[0], 1):4.1f}  {src[ln].strip()[:100]}")
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
