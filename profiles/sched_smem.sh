run() {
  timeout 300 python bench.py --no-cpu-baseline --steps 24 --warmup 3 \
    > gpurun_out/s2_$1.json 2> gpurun_out/s2_$1.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/s2_$1.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernels_ms"]
    print("$1 value=%.0f ms/step=%.3f e2e=%.0f lin_ms=%.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], k["k_newton_linear_ent"][1]))
except Exception as e:
    print("$1 FAILED", e)
P
}
run default
for pct in 58 72 86 100; do RBS_ENT_SMEM_SCHED=2 RBS_ENT_CARVEOUT=$pct run sched2_carveout$pct; done
RBS_ENT_SMEM_SCHED=2 run sched2_rule
