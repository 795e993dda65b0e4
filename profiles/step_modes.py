"""Tuning aid: device-resident windows vs one call per step (no transfers) on the bench workload."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from ribasim_b200.ensemble import perturbed_forcing  # noqa: E402
from ribasim_b200.model import Model  # noqa: E402

B = 4096
p, flat, sym, _ = bench.build_workload(500)
model = Model(p, n_members=B, dt=bench.DT, flat=flat, sym=sym, full_newton=True)
h = model.h
t = 0.0
step = 0


def forcing(day):
    base, fb = bench.base_forcing(p, day)
    for k, v in perturbed_forcing(base, fb, range(B), day).items():
        h.set_member(k, v)


for s in range(48):
    if s % 24 == 0:
        forcing(s // 24)
    h.step(t, bench.DT, want_status=False)
    t += bench.DT
forcing(2)
for mode in ("window24", "single", "window24", "single", "window4"):
    h.synchronize()
    s0 = h.stats()
    t0 = time.perf_counter()
    if mode == "window24":
        h.advance(t, bench.DT, 24, want_status=False)
    elif mode == "window4":
        for k in range(6):
            h.advance(t + 4 * k * bench.DT, bench.DT, 4, want_status=False)
    else:
        for k in range(24):
            h.step(t + k * bench.DT, bench.DT, want_status=False)
    h.synchronize()
    dt = time.perf_counter() - t0
    t += 24 * bench.DT
    s1 = h.stats()
    print(mode, f"{dt / 24 * 1e3:.3f} ms/step", {k: s1[k] - s0[k] for k in ("passes", "newton_iters", "substeps", "rejects")}, flush=True)
model.finalize()
