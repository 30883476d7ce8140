"""Top stall sites from `ncu -i X.ncu-rep --page source --csv` (SASS view) for one kernel instance."""
import csv
import sys


def main(path, kernel_substr, top=22):
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name" and kernel_substr in rows[i][1]:
            hdr = rows[i + 1]
            j = i + 2
            data = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                data.append(rows[j])
                j += 1
            col = {h: k for k, h in enumerate(hdr)}
            stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            tot = sum(int(r[col["Warp Stall Sampling (All Samples)"]] or 0) for r in data if len(r) > 5)
            print(f"kernel {rows[i][1][:70]}  samples {tot}")
            agg = {h: sum(int(r[col[h]] or 0) for r in data if len(r) > 5) for h in stall_cols}
            print("  by reason:", ", ".join(f"{h[6:]}={v / max(tot, 1):.0%}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
            best = sorted((r for r in data if len(r) > 5), key=lambda r: -int(r[col["Warp Stall Sampling (All Samples)"]] or 0))[:top]
            for r in best:
                s = int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
                why = max(stall_cols, key=lambda h: int(r[col[h]] or 0))
                print(f"  {s:6d} {s / max(tot, 1):5.1%} {why[6:]:10s} {r[col['Source']][:90]}")
            return
        i += 1
    print("kernel not found")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 22)
