#!/bin/bash
# Newton linear solve on the bench workload: plain vs compacted shared-memory slots, and the
# shared-memory carve-out (percent of 228 KB; the rest of the 256 KB is L1).  Run under gpurun.
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --no-cpu-baseline --steps 24 --warmup 3 \
    > gpurun_out/compact_$1.json 2> gpurun_out/compact_$1.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/compact_$1.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernels_ms"]
    print("$1 value=%.0f ms/step=%.3f e2e=%.0f lin_ms=%.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], k["k_newton_linear_ent"][1]))
except Exception as e:
    print("$1 FAILED", e)
P
}
run default
RBS_ENT_COMPACT=0 RBS_ENT_CARVEOUT=-1 run plain_driver_carveout
RBS_ENT_COMPACT=0 run plain_rule_carveout
RBS_ENT_CARVEOUT=-1 run compact_driver_carveout
RBS_ENT_CARVEOUT=72 run compact_carveout72
run default_again
