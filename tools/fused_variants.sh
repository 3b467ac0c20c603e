#!/bin/bash
# usage: variants.sh "<EXTRA flags>" ...  -- rebuilds fused.cu with each flag set and benches C2
for v in "$@"; do
  touch mlvfs_b200/csrc/fused.cu
  make -s -C mlvfs_b200/csrc EXTRA="$v" 2>&1 | grep -i "error"
  echo "== $v"
  python -m pytest tests/test_gpu_single_iso.py -x -q -k "wide" 2>&1 | tail -1
  python bench.py --workload C2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['launch_ms'])"
done
