set -x
python profiles/run_profile.py --steps 4 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python profiles/run_profile.py --steps 4 > gpurun_out/ncu_launch.log 2>&1
for k in k_newton_linear_ent k_links k_basin k_assemble_linear k_newton_update k_step_apply; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r1_$k python profiles/run_profile.py --steps 4 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
