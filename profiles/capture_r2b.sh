#!/bin/bash
# ncu evidence of the member-per-lane solve (end of round 2), run on one B200 with
# `gpurun -- 'bash profiles/capture_r2b.sh'`: the launch list of the same command as capture_r2.sh
# (one member group, so full launches take the lane kernel) and full captures of the two kernels
# of the solve group that changed.  A plain run first.
set -x
CMD="python bench.py --ncu-range --steps 4 --spinup 48 --no-extra --no-cpu-baseline --no-profile --groups 1"
$CMD > gpurun_out/r2b_plain.log 2>&1 || { tail -5 gpurun_out/r2b_plain.log; exit 1; }
tail -1 gpurun_out/r2b_plain.log | cut -c1-300
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launch_r2.log 2>&1
for k in k_newton_linear_lane k_assemble_linear; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f \
      -o /tmp/r2_$k $CMD > gpurun_out/ncu_r2_$k.log 2>&1
  ncu -i /tmp/r2_$k.ncu-rep --page raw --csv > gpurun_out/r2_$k.raw.csv 2>/dev/null
  rm -f /tmp/r2_$k.ncu-rep
done
ls -la gpurun_out/ | tail -8
