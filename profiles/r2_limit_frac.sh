# finalize half delayed until the members that wait for it are a fraction of the group's unfinished
# members (RBS_LIMIT_FRAC): parity under the knob, then the whole step
RBS_LIMIT_FRAC=0.5 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size_ensemble or advance_window or forced_window" 2>&1 | tail -2
for f in 0 0.3 0.5 0.7 1.0; do
RBS_LIMIT_FRAC=$f python bench.py --no-cpu-baseline --no-extra --no-profile > gpurun_out/r3f_$f.json 2>gpurun_out/r3f_$f.err
python -c "
import json; d=json.load(open('gpurun_out/r3f_$f.json')); print('limit_frac $f', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), 'passes', d['passes_per_step'], 'launches', d['gpu_launches'], 'min_storage', d['min_storage'])"
done
RBS_LIMIT_FRAC=0.5 python bench.py --no-cpu-baseline --no-extra --no-profile --members 512 > gpurun_out/r3f_512.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r3f_512.json')); print('512 members limit_frac 0.5', round(d['value']), round(d['ms_per_step'],4), 'passes', d['passes_per_step'])"
