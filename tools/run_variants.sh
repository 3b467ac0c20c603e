#!/bin/bash
# usage (on the GPU box): tools/run_variants.sh name ...   -- for every prebuilt build/variants/<name> runs the wide
# kernel's parity tests and the C2 bench line with that library in place; the default library is put back at the end.
cd "$(dirname "$0")/.."
cp mlvfs_b200/libmlvfs_b200.so /tmp/libmlvfs_b200.default.so
for name in "$@"; do
  cp build/variants/$name/libmlvfs_b200.so mlvfs_b200/libmlvfs_b200.so
  echo "== $name: $(cat build/variants/$name/flags.txt) ${VARIANT_ENV}"
  python -m pytest tests/test_gpu_single_iso.py -x -q -k "wide" 2>&1 | tail -1
  python bench.py --workload C2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['launch_ms'])"
done
cp /tmp/libmlvfs_b200.default.so mlvfs_b200/libmlvfs_b200.so
