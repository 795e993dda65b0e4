# ensembles larger than BASELINE's 4096 members on one GPU: launches of the four member groups are
# then >= 2048 members and take the member-per-lane solve (RBS_LIN_LANE=0: shared-memory kernel only)
for m in 8192 16384; do for w in 32 0; do
RBS_LIN_LANE=$w python bench.py --no-cpu-baseline --no-extra --no-profile --members $m > gpurun_out/r3big_${m}_$w.json 2>gpurun_out/r3big_${m}_$w.err
python -c "
import json; d=json.load(open('gpurun_out/r3big_${m}_$w.json')); print('members $m lane $w', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"
done; done
