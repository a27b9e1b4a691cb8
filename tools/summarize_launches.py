"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
into a per-kernel table (markdown): share of device time, launches, and -- when the DRAM counters were collected --
the DRAM traffic and the bandwidth it implies."""
import collections, csv, re, sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path) if l.startswith('"')]
agg = collections.defaultdict(lambda: {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
have_bytes = False
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:90]
    m, unit = row["Metric Name"], row["Metric Unit"]
    a = agg[name]
    if m == "gpu__time_duration.sum":
        a["us"] += {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)
        a["n"] += 1
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        have_bytes = True
        b = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        a["rd" if "read" in m else "wr"] += b
tot = sum(a["us"] for a in agg.values())
n = sum(a["n"] for a in agg.values())
print(f"# {title}\n")
print(f"{n} launches captured, {tot/1e3:.2f} ms of device time (ncu serialises launches and runs them cold-cache: "
      f"compare SHARES, not absolutes).\n")
if have_bytes:
    print("| share | device time (us) | launches | DRAM read (MB) | DRAM write (MB) | DRAM GB/s | kernel |\n|---:|---:|---:|---:|---:|---:|---|")
else:
    print("| share | device time (us) | launches | kernel |\n|---:|---:|---:|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:32]:
    if have_bytes:
        bw = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0
        print(f"| {100*a['us']/tot:.1f}% | {a['us']:.0f} | {a['n']} | {a['rd']/1e6:.0f} | {a['wr']/1e6:.0f} | {bw:.0f} | `{k}` |")
    else:
        print(f"| {100*a['us']/tot:.1f}% | {a['us']:.0f} | {a['n']} | `{k}` |")
