#!/bin/bash
# C4 frames/s for several AMaZE tile-block shapes (threads per block x blocks per SM)
for cfg in "256 3" "512 2" "256 4" "384 3" "256 5" "192 6" "512 3" "384 2"; do
  set -- $cfg
  r=$(MLVB_AMZ_THREADS=$1 MLVB_AMZ_BLOCKS_PER_SM=$2 timeout 200 python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['ms_per_step']/4,2))")
  echo "threads=$1 blocks/SM=$2 -> fps, ms/frame: $r"
done
