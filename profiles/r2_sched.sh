for cfg in "4096 0" "4096 3" "512 1" "512 3" "1024 3" "2048 3" "2048 0"; do
set -- $cfg
RBS_ENT_SMEM_SCHED=$2 python bench.py --no-cpu-baseline --no-extra --members $1 > gpurun_out/r2j_$1_$2.json 2> gpurun_out/r2j_$1_$2.err || tail -3 gpurun_out/r2j_$1_$2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2j_$1_$2.json"))
k=d["roofline"]["kernels_ms"]
print("$1 sched=$2", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), "solve", k["k_newton_linear_ent"])
PY
done
