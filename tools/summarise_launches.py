"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one step (between two launches of the
step's last kernel) grouped by kernel class.  usage: summarise_launches.py <csv> <last-kernel-substring>"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        us = v / 1000.0 if row["Metric Unit"] in ("nsecond", "ns") else v
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name).replace("sky::", "")
        rows.append((name, us))
    return rows


def main():
    rows = load(sys.argv[1])
    last = sys.argv[2]
    idx = [i for i, r in enumerate(rows) if last in r[0]]
    a, b = idx[-2] + 1, idx[-1] + 1
    agg = collections.OrderedDict()
    for n, us in rows[a:b]:
        e = agg.setdefault(n, [0, 0.0])
        e[0] += 1
        e[1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"one step = launches {a}..{b - 1}: {b - a} launches, {tot / 1000:.2f} ms (serialised, cold cache)")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / 1000:.3f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
