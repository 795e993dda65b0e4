# member-per-lane solve: bitwise test, kernel time, per-level clocks, whole step (lane on every launch)
python -m pytest tests/test_gpu_parity.py -x -q -k "lane_solve" 2>&1 | tail -3
RBS_LIN_LANE_MIN=0 RBS_LIN_LANE=32 python profiles/time_linear.py --members 4096 --actives 256 1024 4096 2>&1 | tail -3
RBS_LIN_LANE_MIN=0 RBS_LIB=profiles/micro/librbs_prof.so RBS_LIN_LANE=32 python profiles/time_linear.py --members 1024 --actives 1024 2>&1 | tail -3
for g in ${GROUPS_SWEEP:-4}; do
RBS_LIN_LANE_MIN=0 RBS_LIN_LANE=32 python bench.py --no-cpu-baseline --no-extra --no-profile --groups $g > gpurun_out/r3g_$g.json 2>gpurun_out/r3g_$g.err
python -c "
import json; d=json.load(open('gpurun_out/r3g_$g.json')); print('lane 32 groups $g', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"
done
