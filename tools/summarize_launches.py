"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family and share of the step."""
import collections
import csv
import sys


def main(path, max_rows=40):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("db200::", "")[:56]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        tot.setdefault(name, [0, 0.0])
        tot[name][0] += 1
        tot[name][1] += v
    T = sum(v[1] for v in tot.values())
    print(f"# {path}: {len(rows)} launches, {T / 1e3:.3f} ms total (ncu-serialised, cold cache: compare SHARES)")
    print(f"{'us':>12} {'share':>7} {'n':>5}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:max_rows]:
        print(f"{v[1]:12.1f} {100 * v[1] / T:6.1f}% {v[0]:5d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
