python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for m in 512 1024 4096; do for t in 1 0; do
RBS_HOST_THREADS=$t python bench.py --no-cpu-baseline --no-extra --no-profile --members $m > gpurun_out/r2u_${m}_$t.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2u_${m}_$t.json')); print('members $m host threads $t', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"
done; done
RBS_HOST_THREADS=1 python bench.py --no-cpu-baseline --no-extra --no-profile --members 512 --groups 4 > gpurun_out/r2u_512g4.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2u_512g4.json')); print('512 g4 threads', round(d['value']), round(d['ms_per_step'],4))"
RBS_HOST_THREADS=1 python bench.py --no-cpu-baseline --no-extra --no-profile --members 4096 --groups 8 > gpurun_out/r2u_4096g8.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2u_4096g8.json')); print('4096 g8 threads', round(d['value']), round(d['ms_per_step'],4))"
