"""Round 2: the Newton linear solve with M members per thread against the 2-member tiles
(`rbs_time_linear`: assemble + solve, or the solve alone with RBS_TIME_SOLVE_ONLY=1)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, str(ROOT))
    import bench
    from ribasim_b200.model import Model

    p, flat, sym, _ = bench.build_workload(500)
    model = Model(p, n_members=4096, dt=bench.DT, flat=flat, sym=sym, full_newton=True)
    h = model.h
    t = 0.0
    for s in range(3):
        h.step(t, bench.DT, want_status=False)
        t += bench.DT
    for na in (512, 1024, 4096):
        print(f"LIN_M={os.environ.get('RBS_LIN_M')} SCHED={os.environ.get('RBS_LIN_M_SCHED')} "
              f"solve_only={os.environ.get('RBS_TIME_SOLVE_ONLY')} n_active={na}: "
              f"{1e3 * h.time_linear(na, 20):.1f} us", flush=True)
    model.finalize()
else:
    for m, sched in (("0", "0"), ("4", "0"), ("4", "1"), ("8", "0"), ("8", "1")):
        for so in ("", "1"):
            env = dict(os.environ, RBS_LIN_M=m, RBS_LIN_M_SCHED=sched)
            if so:
                env["RBS_TIME_SOLVE_ONLY"] = "1"
            subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
