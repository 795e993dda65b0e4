#!/bin/bash
# ncu evidence of round 2, run on one B200 with `gpurun -- 'bash profiles/capture_r2.sh'`.
# The profiled command is bench.py itself: the device-resident timed leg of the headline workload
# (HWS-like, 500 basins, 4096 members, one member group so that launches do not overlap) between
# cudaProfilerStart/Stop, after a shortened spin-up.  A plain run first (a number printed under
# ncu is never a bench value).  The .ncu-rep files are turned into CSV on the box and dropped
# (gpurun brings back at most 64 MiB).
set -x
CMD="python bench.py --ncu-range --steps 4 --spinup 48 --no-extra --no-cpu-baseline --no-profile --groups 1"
$CMD > gpurun_out/r2_plain.log 2>&1 || { tail -5 gpurun_out/r2_plain.log; exit 1; }
tail -1 gpurun_out/r2_plain.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launch_r2.log 2>&1
for k in k_newton_linear_ent k_links k_basin k_assemble_linear k_newton_update k_step_apply k_step_limit; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f \
      -o /tmp/r2_$k $CMD > gpurun_out/ncu_r2_$k.log 2>&1
  ncu -i /tmp/r2_$k.ncu-rep --page raw --csv > gpurun_out/r2_$k.raw.csv 2>/dev/null
  rm -f /tmp/r2_$k.ncu-rep
done
# the 100k tree's launch list (config 5), short spin-up
TREE="python bench.py --ncu-range --workload tree --steps 1 --spinup 3 --warmup 1 --no-extra --no-cpu-baseline --no-profile --groups 1"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r2tree.csv $TREE > gpurun_out/ncu_launch_r2tree.log 2>&1
ls -la gpurun_out/ | tail -20
du -sh gpurun_out
