# one ncu --set full capture of a full 4096-member launch of the member-per-lane solve
RBS_LIN_LANE=32 ncu --set full --clock-control none --import-source on -k regex:k_newton_linear_lane -s 2 -c 1 -f -o gpurun_out/r3_lane python profiles/time_linear.py --members 4096 --actives 4096 --reps 1 --groups 1 > gpurun_out/r3_ncu.log 2>&1
tail -3 gpurun_out/r3_ncu.log
