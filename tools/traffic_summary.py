"""Per-kernel DRAM traffic from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
log of ONE eager training step (the last `--last N` launches of each kernel name are kept: earlier ones are warm-up).
usage: traffic_summary.py in.csv out.json [--last-step-launches K]"""
import collections, csv, json, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = collections.OrderedDict()
for row in csv.DictReader(lines):
    key = (row["ID"], re.sub(r"\(.*", "", row["Kernel Name"])[:80])
    d = per.setdefault(key, {})
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
    d[row["Metric Name"]] = v * scale
steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 1
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (_, name), d in per.items():
    a = agg[name]
    a[0] += 1; a[1] += d.get("dram__bytes_read.sum", 0.0); a[2] += d.get("dram__bytes_write.sum", 0.0); a[3] += d.get("gpu__time_duration.sum", 0.0)
out = {k: dict(launches=v[0], dram_read_bytes_per_launch=v[1] / v[0], dram_write_bytes_per_launch=v[2] / v[0],
               dram_bytes_per_launch=(v[1] + v[2]) / v[0], avg_us_under_ncu=v[3] / v[0]) for k, v in agg.items()}
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, v in sorted(out.items(), key=lambda kv: -kv[1]["dram_bytes_per_launch"] * kv[1]["launches"])[:25]:
    print(f"{v['launches']:6d} x {v['dram_bytes_per_launch']/1e6:9.2f} MB  {v['avg_us_under_ncu']:8.1f} us  {k}")
