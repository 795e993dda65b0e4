python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for cfg in "4096 1" "4096 0" "512 1" "512 0"; do
set -- $cfg
RBS_PDL=$2 python bench.py --no-cpu-baseline --no-extra --no-profile --members $1 > gpurun_out/r2k_$1_$2.json 2> gpurun_out/r2k_$1_$2.err || tail -3 gpurun_out/r2k_$1_$2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2k_$1_$2.json"))
print("$1 pdl=$2", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4))
PY
done
RBS_PDL=1 python bench.py --no-cpu-baseline --no-extra --no-profile --workload tree > gpurun_out/r2k_tree1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2k_tree1.json')); print('tree pdl=1', d['value'], d['ms_per_step'], d['newton_iters_per_member_step'])"
RBS_PDL=0 python bench.py --no-cpu-baseline --no-extra --no-profile --workload tree > gpurun_out/r2k_tree0.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2k_tree0.json')); print('tree pdl=0', d['value'], d['ms_per_step'], d['newton_iters_per_member_step'])"
