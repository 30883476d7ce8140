"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path, skip=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    d = collections.defaultdict(list)
    for n, row in enumerate(csv.DictReader(lines)):
        if n < skip:
            continue
        name = row["Kernel Name"].split("(")[0].split("::")[-1]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit.startswith("n") else v * 1000 if unit.startswith("m") else v
        d[name].append(v)
    tot = sum(sum(v) for v in d.values())
    print(f"# {path}: launches {sum(len(v) for v in d.values())} (skipped first {skip}), total {tot:.1f} us")
    print(f"{'kernel':38s} {'n':>5s} {'total_us':>10s} {'share':>6s} {'median':>8s} {'p90':>8s} {'max':>8s}")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        s = sorted(v)
        print(f"{k:38s} {len(v):5d} {sum(v):10.1f} {sum(v) / tot:6.1%} {s[len(s) // 2]:8.2f} {s[int(len(s) * .9)]:8.2f} {s[-1]:8.2f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
