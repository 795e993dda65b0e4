# member-per-lane solve (rbs_lane.cuh) against the shared-memory entry kernel: bitwise test, kernel
# time (assemble + solve per launch, solve alone), per-level clocks of block 0, whole step
python -m pytest tests/test_gpu_parity.py -x -q -k "lane_solve" 2>&1 | tail -3
for w in ${LANES:-0 32 16}; do
RBS_LIN_LANE=$w python profiles/time_linear.py --members 4096 --actives 256 1024 4096 2>&1 | tail -3
done
RBS_LIB=profiles/micro/librbs_prof.so RBS_LIN_LANE=32 python profiles/time_linear.py --members 1024 --actives 1024 2>&1 | tail -3
for w in ${BENCH_LANES:-32 16}; do
RBS_LIN_LANE=$w python bench.py --no-cpu-baseline --no-extra > gpurun_out/r3_$w.json 2>gpurun_out/r3_$w.err
python -c "
import json; d=json.load(open('gpurun_out/r3_$w.json')); print('lane $w', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), 'its', round(d['newton_iters_per_member_step'],4)); print(d['roofline']['kernels_ms'])"
done
