"""Workload used for the ncu captures under profiles/: the bench network (HWS-like, 500 basins,
4096 members), a short spin-up, then a few implicit-Euler steps.  Run as

    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file L.csv \
        python profiles/run_profile.py --steps 2
    ncu --set full --clock-control none --import-source on -k regex:k_newton_linear -s 6 -c 1 \
        -o REP python profiles/run_profile.py --steps 2
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--spinup", type=int, default=4)
ap.add_argument("--members", type=int, default=4096)
ap.add_argument("--basins", type=int, default=500)
ap.add_argument("--full-newton", type=int, default=1)
ap.add_argument("--smem-tile", type=int, default=-1)
args = ap.parse_args()

import bench  # noqa: E402
from ribasim_b200.ensemble import perturbed_forcing  # noqa: E402
from ribasim_b200.model import Model  # noqa: E402

p, flat, sym = bench.build_workload(args.basins)
model = Model(p, n_members=args.members, dt=bench.DT, flat=flat, sym=sym,
              full_newton=bool(args.full_newton), smem_tile=args.smem_tile)
h = model.h
base, fb = bench.base_forcing(p, 0)
f = perturbed_forcing(base, fb, range(args.members), 0)
for k, v in f.items():
    h.set_member(k, v)
t = 0.0
for s in range(args.spinup + args.steps):
    h.step(t, bench.DT, want_status=False)
    t += bench.DT
print("stats", h.stats(), "launches", h.kernel_launches())
model.finalize()
