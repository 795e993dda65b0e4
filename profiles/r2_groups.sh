for g in 2 3 4 6 8; do
python bench.py --no-cpu-baseline --no-extra --no-profile --groups $g > gpurun_out/r2o_g$g.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2o_g$g.json')); print('groups $g', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"
done
for g in 1 2 4; do
python bench.py --no-cpu-baseline --no-extra --no-profile --members 512 --groups $g > gpurun_out/r2o_512g$g.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2o_512g$g.json')); print('512 groups $g', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"
done
