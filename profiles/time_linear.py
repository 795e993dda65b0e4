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
ap.add_argument("--actives", type=int, nargs="*", default=[32, 256, 1024, 4096])
args = ap.parse_args()
import bench  # noqa: E402
from ribasim_b200.model import Model  # noqa: E402

p, flat, sym = bench.build_workload(args.basins)
model = Model(p, n_members=args.members, dt=bench.DT, flat=flat, sym=sym, full_newton=True,
              smem_tile=args.smem_tile, smem_threads=args.smem_threads, lin_mode=args.lin_mode)
h = model.h
t = 0.0
for s in range(3):
    h.step(t, bench.DT, want_status=False)
    t += bench.DT
for na in args.actives:
    if na <= args.members:
        print(f"mode={args.lin_mode} tile={args.smem_tile} threads={args.smem_threads} n_active={na}: "
              f"{1e3 * h.time_linear(na, args.reps):.1f} us", flush=True)
model.finalize()
