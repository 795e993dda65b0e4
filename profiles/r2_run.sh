#!/bin/bash
# usage: r2_run.sh <tag> [pytest]  — GPU tests (optional) + bench at 4096 and 512 members
tag=$1
mkdir -p gpurun_out
if [ "$2" = "pytest" ]; then python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; tail -5 gpurun_out/${tag}_pytest.log; fi
run() { name=$1; shift; env "$@" python bench.py --no-cpu-baseline ${ARGS} > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err || tail -5 gpurun_out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), "passes", d["passes_per_step"], d["roofline"]["kernels_ms"])
except Exception as e: print("$name failed", e)
PY
}
ARGS="" run m4096 X=1
ARGS="--members 512" run m512 X=1
ARGS="--members 512 --groups 4" run m512g4 X=1
ARGS="--groups 8" run m4096g8 X=1
ARGS="" run m4096d3 RBS_PASS_DEPTH=3
ARGS="" run m4096d1 RBS_PASS_DEPTH=1
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cond_graph profiles/micro/cond_graph.cu && /tmp/cond_graph
