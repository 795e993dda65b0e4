# Newton iterations launched per pass (RBS_ITERS_PER_PASS) at several ensemble sizes
for m in 512 1024 4096; do for k in 1 2 3 4; do
RBS_ITERS_PER_PASS=$k python bench.py --no-cpu-baseline --no-extra --no-profile --members $m > gpurun_out/r2t_${m}_$k.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2t_${m}_$k.json')); print('members $m iters/pass $k', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), 'passes', round(d['passes_per_step'],2), 'its', round(d['newton_iters_per_member_step'],4))"
done; done
