for m in 1024 2048; do for g in 2 3 4; do
python bench.py --no-cpu-baseline --no-extra --no-profile --members $m --groups $g > gpurun_out/r2y_${m}_$g.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2y_${m}_$g.json')); print('members $m groups $g', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"
done; done
