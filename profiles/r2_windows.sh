for K in 24 96 192; do
python bench.py --no-cpu-baseline --no-extra --no-profile --steps $K --e2e-window $K > gpurun_out/r2i_$K.json 2> gpurun_out/r2i_$K.err || tail -3 gpurun_out/r2i_$K.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2i_$K.json"))
print($K, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), "passes", round(d["passes_per_step"],2), d["e2e"]["steps_per_call"], round(d["e2e"]["ms_per_step"],4), d["newton_iters_per_member_step"], d["substeps_per_member_step"])
PY
done
