#!/bin/bash
# the two basin kernels of the finalize half (storage / level at the proposed and at the limited
# state), same command as capture_r2.sh: completes the ncu DRAM traffic of the limiter / accept group.
# k_basin launches of the profiled window in order: <1,1> <1,1> <0,1> <0,2> ... (profiles/launches_r2.txt):
# the third and fourth are the first finalize half (2001 of the 4096 members).
CMD="python bench.py --ncu-range --steps 4 --spinup 48 --no-extra --no-cpu-baseline --no-profile --groups 1"
for s in 2 3; do
  tag=k_basin_limit_$([ $s = 2 ] && echo pre || echo post)
  ncu --profile-from-start off --set full --clock-control none --import-source on -k k_basin -s $s -c 1 -f \
      -o /tmp/r2_$tag $CMD > gpurun_out/ncu_r2_$tag.log 2>&1
  ncu -i /tmp/r2_$tag.ncu-rep --page raw --csv > gpurun_out/r2_$tag.raw.csv 2>/dev/null
  rm -f /tmp/r2_$tag.ncu-rep
done
ls -la gpurun_out/ | tail -4
