"""Summarise an .ncu-rep: key metrics per kernel + top stall lines.  usage: python tools/ncu_summary.py rep [nlines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "launch__grid_size", "smsp__inst_executed.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("=====", r[ki][:110])
    for i, h in enumerate(hdr):
        if h in want:
            print(f"   {h:75s} {r[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = [r[1][:90]]; blocks.append(cur); continue
    if cur is not None: cur.append(r)
seen = set()
for b in blocks:
    name, hdr2, data = b[0], b[1], b[2:]
    if name in seen: continue
    seen.add(name)
    si, ii, sc = hdr2.index("# Samples"), hdr2.index("Instructions Executed"), hdr2.index("Source")
    stall_cols = [i for i, h in enumerate(hdr2) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in data if r[si].isdigit()) or 1
    agg = {}
    for r in data:
        for i in stall_cols:
            if r[i].isdigit(): agg[hdr2[i][6:]] = agg.get(hdr2[i][6:], 0) + int(r[i])
    print("\n=== stalls:", name, "samples", tot, "instrs(SASS lines)", len(data))
    print("   by reason:", ", ".join(f"{k}={100*v/tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    for r in sorted(data, key=lambda r: -int(r[si]) if r[si].isdigit() else 0)[:nl]:
        st = sorted(((hdr2[i][6:], int(r[i])) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), key=lambda kv: -kv[1])[:2]
        print(f"   {100*int(r[si])/tot:5.1f}%  x{r[ii]:>8s}  {r[sc][:64]:64s} {st}")
