"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: launch_summary.py in.csv [out.csv] [note]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
    k = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    agg[k][0] += 1; agg[k][1] += v; n += 1
tot = sum(v[1] for v in agg.values())
out = [f"# launches={n} total_ms={tot/1e6:.3f}", "share_pct,launches,avg_us,total_ms,kernel"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{v[1]/tot*100:.2f},{v[0]},{v[1]/v[0]/1e3:.1f},{v[1]/1e6:.3f},\"{k}\"")
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        for note in sys.argv[3:]:
            f.write("# " + note + "\n")
        f.write("\n".join(out) + "\n")
print("\n".join(out[:40]))
