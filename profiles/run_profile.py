"""Workload used for the ncu captures under profiles/: the bench network (HWS-like, 500 basins,
4096 members), the bench's spin-up (untimed, outside the profiled range), then a window of
implicit-Euler steps inside cudaProfilerStart/Stop.  Run as

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file L.csv python profiles/run_profile.py --steps 4
    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:k_newton_linear_smem -c 1 -o REP python profiles/run_profile.py --steps 4
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--spinup", type=int, default=48)
ap.add_argument("--groups", type=int, default=1)
ap.add_argument("--members", type=int, default=4096)
ap.add_argument("--basins", type=int, default=500)
ap.add_argument("--full-newton", type=int, default=1)
ap.add_argument("--smem-tile", type=int, default=-1)
args = ap.parse_args()

import bench  # noqa: E402
from ribasim_b200.ensemble import perturbed_forcing  # noqa: E402
from ribasim_b200.model import Model  # noqa: E402

p, flat, sym, _ = bench.build_workload(args.basins)
model = Model(p, n_members=args.members, dt=bench.DT, flat=flat, sym=sym,
              full_newton=bool(args.full_newton), smem_tile=args.smem_tile,
              n_groups=args.groups)
h = model.h
base, fb = bench.base_forcing(p, 0)
f = perturbed_forcing(base, fb, range(args.members), 0)
for k, v in f.items():
    h.set_member(k, v)
import torch  # noqa: E402

t = 0.0
for s in range(args.spinup):
    if s % 24 == 0:
        base, fb = bench.base_forcing(p, s // 24)
        for k, v in perturbed_forcing(base, fb, range(args.members), s // 24).items():
            h.set_member(k, v)
    h.step(t, bench.DT, want_status=False)
    t += bench.DT
h.synchronize()
s0 = h.stats()
torch.cuda.profiler.start()
h.advance(t, bench.DT, args.steps, want_status=False)
h.synchronize()
torch.cuda.profiler.stop()
s1 = h.stats()
print("profiled window:", {k: s1[k] - s0[k] for k in s1}, "launches", h.kernel_launches())
model.finalize()
