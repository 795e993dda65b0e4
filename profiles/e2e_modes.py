"""Tuning aid: where the end-to-end step time goes (compute, H2D, D2H, overlap)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from ribasim_b200.ensemble import perturbed_forcing  # noqa: E402
from ribasim_b200.model import Model  # noqa: E402

B = 4096
p, flat, sym, _ = bench.build_workload(500)
nb, nfb = flat["n_basin"], flat["n_fb"]
model = Model(p, n_members=B, dt=bench.DT, flat=flat, sym=sym, full_newton=True)
h = model.h
keys = ("precipitation", "potential_evaporation", "drainage", "fb_rate")
base, fb = bench.base_forcing(p, 0)
f = perturbed_forcing(base, fb, range(B), 0)
stage = {k: torch.from_numpy(f[k].copy()).pin_memory() for k in keys}
result = torch.empty((nb, B), dtype=torch.float64).pin_memory()
for k in keys:
    h.set_member(k, f[k])
t = 0.0
for s in range(48):
    h.step(t, bench.DT, want_status=False)
    t += bench.DT
for mode in ("none", "h2d", "d2h", "both", "none"):
    h.synchronize()
    t0 = time.perf_counter()
    for k in range(24):
        if mode in ("h2d", "both", "copies_only"):
            for key in keys:
                h.set_member_ptr_async(key, stage[key].data_ptr(), stage[key].shape[0])
        if mode != "copies_only":
            h.step(t, bench.DT, want_status=False)
            t += bench.DT
        else:
            h.synchronize()
        if mode in ("d2h", "both", "copies_only"):
            h.get_member_ptr_async("storage", result.data_ptr(), nb)
    h.synchronize()
    print(mode, f"{(time.perf_counter() - t0) / 24 * 1e3:.3f} ms/step", flush=True)
model.finalize()
