"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name: total ms, launches, share."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0.0, 0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        agg[name][0] += ms
        agg[name][1] += 1
        tot += ms
    print(f"total {tot:.2f} ms in {sum(v[1] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]:9.3f} ms  x{v[1]:5d}  {100 * v[0] / tot:5.1f} %  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
