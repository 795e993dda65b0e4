"""Tuning aid: the linear-solve tiers must give the same trajectories on the bench network with
perturbed (distinct) members."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from ribasim_b200.ensemble import perturbed_forcing  # noqa: E402
from ribasim_b200.model import Model  # noqa: E402

B = 72
p, flat, sym, _ = bench.build_workload(500)
base, fb = bench.base_forcing(p, 0)
forcing = perturbed_forcing(base, fb, range(B), 0)
res = {}
for lm in (2, 3):
    model = Model(p, n_members=B, dt=bench.DT, flat=flat, sym=sym, full_newton=True, lin_mode=lm, n_groups=1)
    h = model.h
    for k, v in forcing.items():
        h.set_member(k, v)
    t = 0.0
    try:
        for s in range(3):
            st = h.step(t, bench.DT)
            t += bench.DT
        res[lm] = h.get_member("u", flat["n_state"]).copy()
        print(lm, "ok", h.stats())
    except Exception as e:  # noqa: BLE001
        print(lm, "FAILED", e, h.stats())
    model.finalize()
if 2 in res and 3 in res:
    sc = 1e-5 + 1e-5 * np.abs(res[2])
    print("max scaled diff", np.max(np.abs(res[2] - res[3]) / sc))
