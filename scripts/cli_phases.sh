#!/bin/bash
# phases of one CLI run on 1e8 + 1e8 lists in /dev/shm
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT"
python - <<'PY'
import sys; sys.path.insert(0, '.')
import bench
from pathlib import Path
d = Path('/dev/shm/gt4cli'); d.mkdir(exist_ok=True)
bench.build_sample_lists(d, 1e8, 0.5, 25)
PY
cd /dev/shm/gt4cli
for i in 1 2; do
  s=$(date +%s.%N)
  "$ROOT/genometester4_b200/gt4gpu-compare" sample_A.list sample_B.list -u -D -o g 2>&1 | grep -E "gt4gpu"
  e=$(date +%s.%N); python -c "print(\"wall\", round($e - $s, 3), \"s\")"
done
rm -rf /dev/shm/gt4cli
