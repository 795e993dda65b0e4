for m in 1024 2048; do for p in 0 1; do
RBS_PDL=$p python bench.py --no-cpu-baseline --no-extra --no-profile --members $m > gpurun_out/r2x_${m}_$p.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2x_${m}_$p.json')); print('members $m pdl $p', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"
done; done
RBS_ENT_SMEM_SCHED=0 python bench.py --no-cpu-baseline --no-extra --no-profile --members 1024 > gpurun_out/r2x_1024_s0.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2x_1024_s0.json')); print('members 1024 sched 0', round(d['value']), round(d['ms_per_step'],4))"
