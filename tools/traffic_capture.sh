#!/bin/bash
# DRAM traffic per kernel of ONE eager training step of bench.py, captured with ncu from THESE kernel sources (the capture
# records bench.lib_hash(), a sha256 over csrc/ and the C header; bench.py only uses it when it matches the sources it runs).
#     gpurun --timeout 1200 -- tools/traffic_capture.sh        -> gpurun_out/r02_traffic_raw.csv, profiles/r02_traffic_train.json
set -u
CMD="python bench.py --no-graph --steps 1 --warmup 2 --no-cpu-baseline --no-instrument --no-secondary --no-parity"
mkdir -p gpurun_out
$CMD > gpurun_out/traffic_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/traffic_plain.log; exit 1; }
# skip the first two steps (lazy initialisation, warm-up), take ONE step's worth of this library's launches (704 per step;
# any 704 consecutive launches are one full step cyclically)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b2c:: -s ${TRAFFIC_SKIP:-1430} -c ${TRAFFIC_COUNT:-704} \
    --csv --log-file gpurun_out/r02_traffic_raw.csv $CMD > gpurun_out/traffic_ncu.log 2>&1
echo "ncu rc=$?"
python tools/traffic_summary.py gpurun_out/r02_traffic_raw.csv gpurun_out/r02_traffic_kernels.json | head -20
python - <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench
k = json.load(open("gpurun_out/r02_traffic_kernels.json"))
h = bench.lib_hash()
json.dump(dict(src_sha256_16=h, command="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:b2c:: -s 1430 -c 704 python bench.py --no-graph --steps 1 --warmup 2 ... (one eager training step)",
               kernels=k), open("gpurun_out/r02_traffic_train.json", "w"), indent=1)
print("kernel sources", h)
PY
