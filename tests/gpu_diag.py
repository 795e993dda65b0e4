"""Diagnostic parity sweep (not a pytest test): prints per-model max errors GPU vs oracle."""
import sys, json, time, traceback
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from tests.helpers import *  # noqa

names = sys.argv[1:] or ["basic", "backwater", "trivial", "linear_resistance", "rating_curve",
                         "manning_resistance", "misc_nodes", "user_demand", "pid_control",
                         "pid_control_equation", "outlet_continuous_control"]
for name in names:
    try:
        B = 7
        case = make_case(name, B, seed=1)
        flat, p = case["flat"], case["p"]
        h = gpu_handle_for(case)
        t = case["t"]
        du_g, cache_g = h.rhs(case["ur"], t, with_cache=True)
        jv, wv = h.jac(case["ur"], t, gamma=3600.0)
        rng = np.random.default_rng(5)
        b = rng.normal(size=(flat["n_reduced"], B))
        x = h.solve(b)
        e = dict(du=0, S=0, lvl=0, J=0, Joff=0, W=0, solve=0)
        for m in range(B):
            om = oracle_for_member(case, m)
            du_o, cache_o = om.rhs(case["ur"][:, m], t, with_cache=True)
            e["du"] = max(e["du"], rel_err(du_g[:, m], du_o))
            e["S"] = max(e["S"], rel_err(cache_g["storage"][:, m], cache_o["storage"]))
            e["lvl"] = max(e["lvl"], rel_err(cache_g["level"][:, m], cache_o["level"]))
            Jo = om.jac_dense(case["ur"][:, m], t)
            Jg = csr_to_dense(flat, jv[:, m])
            sc = np.maximum(np.abs(Jo).max(axis=1, keepdims=True), 1e-300)
            e["J"] = max(e["J"], float(np.nanmax(np.abs(Jg - Jo) / sc)))
            Wo = om.W_inner(csc_vals_from_dense(flat, Jo), 3600.0)
            Wg = w_to_dense(flat, wv[:, m])
            e["W"] = max(e["W"], float(np.nanmax(np.abs(Wg - Wo) / np.maximum(np.abs(Wo).max(), 1e-300))))
            xo = np.linalg.solve(Wg, b[:, m])
            e["solve"] = max(e["solve"], float(np.abs(x[:, m] - xo).max() / np.abs(xo).max()))
        print(name, json.dumps({k: float(f"{v:.3e}") for k, v in e.items()}), flush=True)
        h.close()
    except Exception:
        print(name, "FAILED"); traceback.print_exc()
