"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*.csv) per kernel."""
import collections
import csv
import sys


def summarize(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split("(")[0].replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += float(r[-1].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out = [f"| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k[:60]} | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / tot:.3f} |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
