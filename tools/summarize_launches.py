"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid): count, mean/min/max µs, share."""
import collections
import csv
import sys


def main(path):
    rows = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(rows):
        name = row["Kernel Name"].replace("<unnamed>::", "")
        name = name.split("(")[0][-60:]
        key = (name, row["Grid Size"], row["Block Size"])
        agg.setdefault(key, []).append(float(row["Metric Value"].replace(",", "")) / 1e3)
    total = sum(sum(v) for v in agg.values())
    print("| kernel | grid | block | launches | mean µs | min µs | max µs | share of listed time |")
    print("|---|---|---|---|---|---|---|---|")
    for (name, grid, block), v in agg.items():
        print(f"| `{name}` | {grid} | {block} | {len(v)} | {sum(v)/len(v):.1f} | {min(v):.1f} | {max(v):.1f} | {100*sum(v)/total:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1])
