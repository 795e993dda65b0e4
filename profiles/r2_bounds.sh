# resident-block bounds of k_basin / k_links (register allocation vs occupancy); builds made with
# nvcc ... -DRBS_BASIN_BLOCKS=b -DRBS_LINKS_BLOCKS=l -o ribasim_b200/librbs_b${b}l${l}.so
for v in "" b5l4 b4l4 b6l5 b6l3; do
if [ -z "$v" ]; then lib=""; else lib="$PWD/ribasim_b200/librbs_$v.so"; fi
RBS_LIB=$lib python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2s_$v.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2s_$v.json')); k=d['roofline']['kernels_ms']; print('variant [$v]', round(d['value']), round(d['ms_per_step'],4), 'basin', k['k_basin<jac>'][1], 'links', k['k_links<jac>'][1], 'limit_pre', k['k_basin_limit_pre'][1])"
done
