"""Diagnostic (not a pytest test): print Jacobian mismatches GPU vs oracle for one model."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from tests.helpers import *  # noqa

name = sys.argv[1]
case = make_case(name, 9, seed=1)
flat, t = case["flat"], case["t"]
h = gpu_handle_for(case)
jv, wv = h.jac(case["ur"], t, gamma=3600.0)
for m in range(9):
    om = oracle_for_member(case, m)
    Jo = om.jac_dense(case["ur"][:, m], t)
    Jg = csr_to_dense(flat, jv[:, m])
    rs = np.maximum(np.abs(Jo).max(axis=1, keepdims=True), 1e-300)
    bad = np.argwhere(np.abs(Jg - Jo) / rs > 1e-12)
    for r, c in bad[:6]:
        print(m, r, c, Jg[r, c], Jo[r, c])
print("tab_slope_right", flat["tab_slope_right"], "tab_ptr", flat["tab_ptr"])
