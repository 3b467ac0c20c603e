#!/bin/bash
# usage (on the GPU box): [TESTS="tests/test_gpu_x.py -k expr"] [WORKLOAD=C4] [BENCH_ARGS="--quick"] tools/run_variants.sh name ...
# For every prebuilt build/variants/<name> (tools/build_variants.sh) runs the parity tests and one bench line with that
# library in place; the default library is put back at the end.  Defaults: the wide kernel's tests and the C2 line.
cd "$(dirname "$0")/.."
TESTS=${TESTS:-tests/test_gpu_single_iso.py -k wide}
WORKLOAD=${WORKLOAD:-C2}
cp mlvfs_b200/libmlvfs_b200.so /tmp/libmlvfs_b200.default.so
for name in default "$@"; do
  if [ "$name" != default ]; then cp build/variants/$name/libmlvfs_b200.so mlvfs_b200/libmlvfs_b200.so; fi
  echo "== $name: $(cat build/variants/$name/flags.txt 2>/dev/null) ${VARIANT_ENV}"
  python -m pytest $TESTS -x -q 2>&1 | tail -1
  python bench.py --workload $WORKLOAD --only --no-cpu-baseline $BENCH_ARGS 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d.get('roofline',{})
print('  value', round(d['value'],1), 'sustained', round(d.get('sustained',{}).get('value',0),1), 'frac', r.get('frac'), 'launch_ms', r.get('launch_ms'), 'stages', r.get('stage_ms_per_step', d.get('stage_ms_per_step')))"
done
cp /tmp/libmlvfs_b200.default.so mlvfs_b200/libmlvfs_b200.so
