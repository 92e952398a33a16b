#!/bin/bash
# Runs every GPU kernel test in its own process (a trap / illegal access poisons only that process) and
# collects a summary under gpurun_out/.  Usage (on the GPU box): bash tools/gpu_probe.sh [test-file] [k-filter...]
mkdir -p gpurun_out
FILE=${1:-tests/test_kernels_gpu.py}
shift
if [ $# -gt 0 ]; then TESTS="$@"; else
TESTS=$(python - "$FILE" <<'PY'
import re,sys
print(" ".join(re.findall(r"^def (test_\w+)", open(sys.argv[1]).read(), re.M)))
PY
)
fi
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
: > gpurun_out/probe_summary.txt
for t in $TESTS; do
  timeout ${PROBE_TIMEOUT:-420} python -m pytest "$FILE" -m gpu -q -x -k "$t" -p no:cacheprovider > gpurun_out/probe_$t.log 2>&1
  rc=$?
  echo "$t rc=$rc $(tail -1 gpurun_out/probe_$t.log)" >> gpurun_out/probe_summary.txt
done
cat gpurun_out/probe_summary.txt
