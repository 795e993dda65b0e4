#!/bin/bash
tag=$1
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python bench.py --no-cpu-baseline ${ARGS} > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err || tail -5 gpurun_out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), "passes", round(d["passes_per_step"],2), {k:v[1] for k,v in d["roofline"]["kernels_ms"].items()})
except Exception as e: print("$name failed", e)
PY
}
