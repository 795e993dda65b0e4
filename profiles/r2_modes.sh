#!/bin/bash
# round-2 mode comparison on one B200: per-kernel times of the stepping variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
run() { name=$1; shift; env "$@" python bench.py --no-cpu-baseline ${ARGS} > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["kernels_ms"])
except Exception as e: print("$name failed", e)
PY
}
ARGS="" run r1path RBS_R2_PASS=0 RBS_LIN_M=0
ARGS="" run linm8 RBS_R2_PASS=0 RBS_LIN_M=8
ARGS="" run linm4 RBS_R2_PASS=0 RBS_LIN_M=4
ARGS="" run fin RBS_R2_PASS=1 RBS_LIN_M=0
ARGS="--lin-mode 1" run l2tile RBS_R2_PASS=0 RBS_LIN_M=0
ARGS="--lin-mode 2" run ldu RBS_R2_PASS=0 RBS_LIN_M=0
ARGS="--members 512" run m512 RBS_R2_PASS=0 RBS_LIN_M=0
