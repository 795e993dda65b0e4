#!/bin/bash
# Stream path vs tile window kernel (rbs_tile.cuh) on the bench workload; run under gpurun.
mkdir -p gpurun_out
for tw in 0 2 4; do
  for mem in 4096 512; do
    RBS_TILE_WINDOW=$tw timeout 300 python bench.py --no-cpu-baseline --members $mem --steps 24 --warmup 3 \
      > gpurun_out/tile_${tw}_${mem}.json 2> gpurun_out/tile_${tw}_${mem}.err
    python - <<P
import json
try:
    d = json.loads(open("gpurun_out/tile_${tw}_${mem}.json").read().strip().splitlines()[-1])
    print("tile_window=$tw members=$mem value=%.0f ms/step=%.3f e2e=%.0f launches=%d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
except Exception as e:
    print("tile_window=$tw members=$mem FAILED", e)
P
  done
done
