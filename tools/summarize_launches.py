"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections, csv, re, sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:100]
    agg[name][0] += 1
    agg[name][1] += v
    n += 1
tot = sum(v for _, v in agg.values())
print(f"# {title}\n")
print(f"{n} launches captured, {tot/1e3:.2f} ms of device time (ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes).\n")
print("| share | device time (us) | launches | kernel |\n|---:|---:|---:|---|")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"| {100*v/tot:.1f}% | {v:.0f} | {c} | `{k}` |")
