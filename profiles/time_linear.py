"""Microbenchmark of the Newton linear-solve kernel on the bench network (tuning aid)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=4096)
ap.add_argument("--basins", type=int, default=500)
ap.add_argument("--smem-tile", type=int, default=-1)
ap.add_argument("--smem-threads", type=int, default=1024)
ap.add_argument("--lin-mode", type=int, default=-1)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--groups", type=int, default=0)
ap.add_argument("--actives", type=int, nargs="*", default=[32, 256, 1024, 4096])
args = ap.parse_args()
import bench  # noqa: E402
from ribasim_b200.model import Model  # noqa: E402

p, flat, sym, _ = bench.build_workload(args.basins)
model = Model(p, n_members=args.members, dt=bench.DT, flat=flat, sym=sym, full_newton=True,
              smem_tile=args.smem_tile, smem_threads=args.smem_threads, lin_mode=args.lin_mode,
              n_groups=args.groups)
h = model.h
t = 0.0
for s in range(3):
    h.step(t, bench.DT, want_status=False)
    t += bench.DT
import os  # noqa: E402

for na in args.actives:
    if na <= args.members:
        os.environ.pop("RBS_TIME_SOLVE_ONLY", None)
        both = 1e3 * h.time_linear(na, args.reps)
        os.environ["RBS_TIME_SOLVE_ONLY"] = "1"  # read per launch: the solve kernel alone
        solve = 1e3 * h.time_linear(na, args.reps)
        print(f"lane={os.environ.get('RBS_LIN_LANE', '0')} mode={args.lin_mode} tile={args.smem_tile} n_active={na}: "
              f"assemble + solve {both:.1f} us, solve {solve:.1f} us", flush=True)
os.environ.pop("RBS_TIME_SOLVE_ONLY", None)
model.finalize()
