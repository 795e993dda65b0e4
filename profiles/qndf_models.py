import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from ribasim_b200.model import Model
from ribasim_b200.network import build_params
from ribasim_b200.testmodels import MODELS
names = ["basic_transient", "pump_discrete_control", "pid_control", "tabulated_rating_curve_control", "flow_condition",
         "level_boundary_condition", "user_demand", "outlet", "backwater", "misc_nodes", "linear_resistance",
         "rating_curve", "manning_resistance", "two_basin", "junction_combined", "drought", "circular_flow",
         "outlet_continuous_control", "discrete_control_of_pid_control", "storage_condition", "minimal_subnetwork"]
for name in names:
    t0 = time.time()
    try:
        m = MODELS[name]()
        if m.solver.get("algorithm", "QNDF") != "QNDF":
            m.solver.update(algorithm="QNDF", dt=None)
        model = Model(build_params(m))
        assert model.adaptive, name
        model.update_until(model.t_end)
        S = model.storage()[:, 0]
        rec = model.control_record()[:6] if model.flat.ints["n_dc"] else None
        print(name, "ok steps", model.qndf.n_steps, "rej", model.qndf.n_rejected, "S", np.round(S[:4], 3), "rec", rec, round(time.time() - t0, 1), "s", flush=True)
        model.finalize()
    except Exception as e:
        print(name, "FAILED", type(e).__name__, str(e)[:200], flush=True)
