for r in 64 32 16; do
for m in 4096 512; do
RBS_ROWS_PER_BLOCK=$r python bench.py --no-cpu-baseline --no-extra --members $m > gpurun_out/r2p_${m}_$r.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2p_${m}_$r.json')); k=d['roofline']['kernels_ms']; print('rows $r members $m', round(d['value']), round(d['ms_per_step'],4), 'update', k['k_newton_update'], 'iters', round(d['newton_iters_per_member_step'],4))"
done; done
