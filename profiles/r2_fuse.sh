python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for cfg in "4096 1" "4096 0" "512 1" "512 0"; do
set -- $cfg
RBS_FUSE=$2 python bench.py --no-cpu-baseline --no-extra --members $1 > gpurun_out/r2m_$1_$2.json 2> gpurun_out/r2m_$1_$2.err || tail -3 gpurun_out/r2m_$1_$2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2m_$1_$2.json"))
print("$1 fuse=$2", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), {k:v[1] for k,v in d["roofline"]["kernels_ms"].items()})
PY
done
