"""Tuning aid: stream path vs cooperative window kernel for single instances and small ensembles."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ribasim_b200.model import Model  # noqa: E402
from ribasim_b200.network import build_params  # noqa: E402
from ribasim_b200.testmodels import MODELS, hws_like_model  # noqa: E402

for name, B in (("backwater", 1), ("backwater", 32), ("hws", 1), ("hws", 32), ("hws", 128)):
    for coop in (0, 4096):
        os.environ["RBS_COOP_MEMBERS"] = str(coop)
        tables = hws_like_model(n_basins=500, n_days=3) if name == "hws" else MODELS[name]()
        m = Model(build_params(tables), n_members=B, dt=3600.0)
        m.update_until(6 * 3600.0)
        m.h.synchronize()
        s0 = m.h.stats()
        t0 = time.perf_counter()
        m.update_until(54 * 3600.0)
        m.h.synchronize()
        dt = time.perf_counter() - t0
        s1 = m.h.stats()
        print(f"{name:9s} B={B:4d} {'coop  ' if coop else 'stream'} {dt / 48 * 1e3:7.3f} ms/step  "
              f"passes/step {(s1['passes'] - s0['passes']) / 48:.1f}", flush=True)
        m.finalize()
