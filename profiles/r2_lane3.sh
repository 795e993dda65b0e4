# lane solve with / without the L1 prefetch of the next entry's operands (RBS_LANE_PF build macro)
python -m pytest tests/test_gpu_parity.py -x -q -k "lane_solve" 2>&1 | tail -2
RBS_LIN_LANE_MIN=0 python profiles/time_linear.py --members 4096 --actives 256 1024 4096 2>&1 | tail -3
RBS_LIB=profiles/micro/librbs_nopf.so RBS_LIN_LANE_MIN=0 python profiles/time_linear.py --members 4096 --actives 256 1024 4096 2>&1 | tail -3
RBS_LIN_LANE_MIN=0 python bench.py --no-cpu-baseline --no-extra --no-profile > gpurun_out/r3pf.json 2>gpurun_out/r3pf.err
python -c "
import json; d=json.load(open('gpurun_out/r3pf.json')); print('lane always, prefetch', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"
