#!/usr/bin/env python
"""Benchmark of the implicit-integration inner loop on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, host cores)

A "step" is one outer implicit-Euler step (dt = 3600 s) of every ensemble member resident on
the rank: all Newton iterations, sub-steps, limit_flow!, the negative-storage check and the
exact-forcing bookkeeping.  Workload (BASELINE.json config 4): HWS-like synthetic network
(the national HWS model is a private DVC artefact, SURVEY intro fact 5) with 4096 perturbed-
forcing members per GPU; members are independent, so ranks share nothing inside the loop and
the only collective is the final gather.  One JSON line is printed by rank 0.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DT = 3600.0
METRIC = "scenario-steps/sec"
UNIT = "scenario-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=4096, help="ensemble members per GPU")
    ap.add_argument("--basins", type=int, default=500)
    ap.add_argument("--workload", default="hws", choices=["hws", "tree"],
                    help="hws: HWS-like network (BASELINE config 4); tree: random tree (config 5)")
    ap.add_argument("--spinup", type=int, default=48, help="untimed steps from the initial state")
    ap.add_argument("--full-newton", type=int, default=1)
    ap.add_argument("--lu-tile", type=int, default=0)
    ap.add_argument("--lu-mode", type=int, default=-1)
    ap.add_argument("--smem-tile", type=int, default=-1)
    ap.add_argument("--smem-threads", type=int, default=1024)
    ap.add_argument("--lin-mode", type=int, default=-1)
    ap.add_argument("--groups", type=int, default=0, help="member groups on own streams (0 = auto)")
    ap.add_argument("--e2e-window", type=int, default=24,
                    help="outer steps per rbs_advance_window call of the end-to-end leg (1 = one call per step)")
    ap.add_argument("--strong", action="store_true", help="keep --members as the job total")
    ap.add_argument("--cpu-members", type=int, default=0, help="sample size of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- workload
def build_workload(n_basins: int, workload: str = "hws"):
    from ribasim_b200.flatten import flatten
    from ribasim_b200.network import build_params
    from ribasim_b200.symbolic import analyse
    from ribasim_b200.testmodels import hws_like_model, random_tree_model

    if workload == "tree":
        tables = random_tree_model(n_basins=n_basins, n_hours=24 * 30)
    else:
        tables = hws_like_model(n_basins=n_basins, n_days=30)
    p = build_params(tables)
    flat = flatten(p)
    sym = analyse(flat["n_reduced"], flat["wrow_ptr"], flat["wcol"])
    return p, flat, sym


def base_forcing(p, day: int):
    """Unperturbed basin vertical fluxes and boundary inflows of calendar day `day`."""
    t = day * 86400.0
    base = {k: np.array([s(t) for s in p.basin.forcing[k]]) for k in
            ("precipitation", "potential_evaporation", "drainage")}
    fb = np.array([s(t) for s in p.nodes["flow_boundary"]["flow_rate"]])
    return base, fb


def algorithmic_bytes(flat, sym) -> dict[str, float]:
    """Per member, per call; shared topology/parameter tables count 0."""
    nb, nc, npid = flat["n_basin"], flat["n_conn"], flat["n_pid"]
    n, nr = flat["n_state"], flat["n_reduced"]
    return dict(
        rhs=48 * nb + 8 * nc + 16 * npid,
        jac=32 * nb + 8 * npid + 8 * flat["nnz_j"],
        # linear solve of one Newton iteration with the factors resident in shared memory:
        # read W (assembled on the fly from J), read b_r, write x
        solve=8 * flat["nnz_w"] + 16 * nr,
        # frozen-Newton mode keeps the factors for later iterations: one write, one read each
        lu_spill=8 * sym.n_lu,
        limiter=24 * n,
    )


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device = device
        self.samples: list[float] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                bits = int(out[2].strip(), 16) if out[2].strip().startswith("0x") else 0
                names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
                         0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
                         0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
                         0x100: "display_clock_setting"}
                for b, nm in names.items():
                    if bits & b and nm != "gpu_idle":
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons))


# ----------------------------------------------------------------------------- CPU arm
def cpu_arm(p, flat, n_members: int, n_steps: int, spinup: int, full_newton: int, threads: int):
    """Oracle port (C++ restatement of the reference path) on host threads: one instance per
    member, same stepping options as the device."""
    from oracle.oracle import OracleBatch, OracleModel, _dp, advance_opts
    from ribasim_b200.ensemble import perturbed_forcing
    from ribasim_b200.timecache import eval_time_params

    proto = OracleModel(p)
    proto.set_pattern(flat.jac_rows_csc, flat.jac_cols_csc)
    proto.set_time_params(eval_time_params(p, 0.0))
    _, c0 = proto.rhs(np.zeros(proto.n_reduced), 0.0, with_cache=True)
    lv = np.ascontiguousarray(c0["level"])
    members = []
    for m in range(n_members):
        om = proto.clone()
        om.set_pattern(flat.jac_rows_csc, flat.jac_cols_csc)
        om.L.orc_set_level_prev(om.h, _dp(lv))
        members.append(om)
    # The CPU arm always runs the reference's own Newton variant (J and W evaluated once per
    # attempt, OrdinaryDiffEq NLNewton) — it is also the faster one on the CPU.
    del full_newton
    batch = OracleBatch(members, advance_opts(1e-5, 1e-5, 10, 0, 4, 20, 1, 1, 2), threads)

    def set_day(day):
        base, fb = base_forcing(p, day)
        f = perturbed_forcing(base, fb, range(n_members), day)
        for j, om in enumerate(members):
            for k in ("precipitation", "potential_evaporation", "drainage"):
                om.set_f64(k, f[k][:, j])
            om.set_f64("fb_rate_step", f["fb_rate"][:, j])
            om.set_f64("fb_rate_now", f["fb_rate"][:, j])

    t = 0.0
    step = 0

    def advance_steps(k):
        nonlocal t, step
        stats = dict(failed=0, substeps=0, iters=0)
        done = 0
        while done < k:
            if step % 24 == 0:
                set_day(step // 24)
            chunk = min(k - done, 24 - step % 24)
            s = batch.run(t, DT, chunk)
            for key in stats:
                stats[key] += s[key]
            t += chunk * DT
            step += chunk
            done += chunk
        return stats

    advance_steps(spinup)
    t0 = time.perf_counter()
    stats = advance_steps(n_steps)
    el = time.perf_counter() - t0
    return n_members * n_steps / el, el, stats


def reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm for this path on the box's host cores.
    Julia is not installed (SURVEY intro fact 4), so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    p, flat, sym = build_workload(args.basins, args.workload)
    sample = args.cpu_members or 32 * cores
    # warm-up steps are part of the (untimed) spin-up of the sample
    val, el, stats = cpu_arm(p, flat, sample, args.steps, args.spinup + args.warmup,
                             args.full_newton, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, flat, sym, sample, 1) | {"sample_members": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} members x {args.steps} steps after "
                                   f"{args.spinup + args.warmup} spin-up steps, one oracle "
                                   f"instance per member on {cores} host threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "newton_iters_per_member_step": stats["iters"] / (sample * args.steps),
        "substeps_per_member_step": stats["substeps"] / (sample * args.steps),
    }
    print(json.dumps(line), flush=True)


def workload_config(args, flat, sym, members_local, world):
    return {
        "workload": f"{'random-tree' if args.workload == 'tree' else 'hws_like'} synthetic network, {flat['n_basin']} basins x "
                    f"{members_local} members per GPU, dt=3600 s implicit Euler",
        "n_basin": flat["n_basin"], "n_state": flat["n_state"], "n_reduced": flat["n_reduced"],
        "nnz_j": flat["nnz_j"], "nnz_w": flat["nnz_w"], "n_lu": sym.n_lu,
        "members_per_gpu": members_local, "members_total": members_local * world,
        "dt_s": DT, "spinup_steps": args.spinup, "full_newton": args.full_newton,
        "l2": "per-step working set (>= 0.5 GB of per-member state) exceeds the 126 MB L2",
        "parallelism": f"members sharded over {world} GPU(s), no collective in the loop",
    }


# ----------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator comes up: keep
        # stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from ribasim_b200.ensemble import gather_members, member_range, perturbed_forcing
    from ribasim_b200.model import Model

    p, flat, sym = build_workload(args.basins, args.workload)
    if args.strong:
        lo, hi = member_range(args.members, rank, world)
        total_members = args.members
    else:
        lo, hi = rank * args.members, (rank + 1) * args.members
        total_members = args.members * world
    nm = hi - lo
    nb, n, nfb = flat["n_basin"], flat["n_state"], flat["n_fb"]
    model = Model(p, n_members=nm, device=local_rank, dt=DT, flat=flat, sym=sym,
                  full_newton=bool(args.full_newton), lu_tile=args.lu_tile, lu_mode=args.lu_mode,
                  smem_tile=args.smem_tile, smem_threads=args.smem_threads, lin_mode=args.lin_mode,
                  n_groups=args.groups)
    h = model.h
    stream = torch.cuda.ExternalStream(h.stream_ptr(), device=local_rank)

    # pinned staging buffers: two days of perturbed forcing + the result read-back
    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64).pin_memory()

    keys = ("precipitation", "potential_evaporation", "drainage", "fb_rate")
    stage: dict[int, dict] = {}  # day -> pinned per-member forcing arrays
    result = pinned((nb, nm))

    def fill_day(day):
        """Host-side generation of one day of perturbed forcing into pinned staging buffers
        (the ensemble's input producer; kept outside every timed region)."""
        if day in stage:
            return
        stage[day] = {k: pinned((nfb if k == "fb_rate" else nb, nm)) for k in keys}
        base, fb = base_forcing(p, day)
        perturbed_forcing(base, fb, range(lo, hi), day, out={k: v.numpy() for k, v in stage[day].items()})

    def upload(day):
        for k in keys:
            h.set_member_ptr(k, stage[day][k].data_ptr(), stage[day][k].shape[0])

    t = 0.0
    step = 0

    W = max(1, min(args.e2e_window, 24))
    results = [pinned((W, nb, nm)) for _ in range(2)]
    e2e_count = {"h2d": 0, "d2h": 0, "calls": 0}

    def stage_window(step0, w):
        """Forcing slices of outer steps step0 .. step0 + w - 1, one upload per outer step from the
        pinned arrays of its day (a window may span a daily change)."""
        for s in range(w):
            day = (step0 + s) // 24
            for key in keys:
                src = stage[day][key]
                h.stage_forcing_window(key, src.data_ptr(), src.shape[0], 1, first_step=s)
                e2e_count["h2d"] += src.shape[0] * nm * 8

    def run_steps(k, e2e=False):
        """Device-resident path: windows of outer steps up to the next daily forcing change in one
        rbs_advance call.  e2e path: forced windows of W outer steps through the C ABI with HOST
        buffers - every outer step's forcing slice is uploaded from pinned memory and every outer
        step's storages are downloaded inside the timed region; the upload of window k+1 and the
        download of window k-1 overlap with the computation of window k."""
        nonlocal t, step
        done = 0
        while done < k:
            if e2e:
                w = min(W, k - done)
                left = k - done - w
                stage_window(step + w, min(W, left) if left > 0 else W)  # live when this window returns
                h.advance_window(t, DT, w, results[e2e_count["calls"] % 2].data_ptr(), want_status=False)
                e2e_count["d2h"] += w * nb * nm * 8
                e2e_count["calls"] += 1
            else:
                w = min(k - done, 24 - step % 24)
                h.advance(t, DT, w, want_status=False)
            t += w * DT
            step += w
            done += w

    # spin-up from the initial state with daily forcing changes (untimed)
    model._member_forcing = True
    for s in range(args.spinup):
        if s % 24 == 0:
            fill_day(s // 24)
            upload(s // 24)
        run_steps(1)
    fill_day(step // 24)
    upload(step // 24)
    run_steps(args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, e2e=False):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        run_steps(k, e2e)
        if e2e:
            h.synchronize()  # the download of the last window's results belongs to the timed region
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    sampler = ClockSampler(local_rank)
    sampler.start()
    s0, l0 = h.stats(), h.kernel_launches()
    ms = timed(args.steps)
    s1, l1 = h.stats(), h.kernel_launches()
    clocks = sampler.stop()
    value = total_members * args.steps / (ms * 1e-3)

    # end to end through the public API with host buffers (H2D forcing + D2H storages per step);
    # the forcing of the days covered is generated beforehand, the copies are timed
    for d in [d for d in stage if d < step // 24]:
        del stage[d]
    for d in range(step // 24, (step + args.steps + 3 * W) // 24 + 1):
        fill_day(d)
    stage_window(step, W)
    h.commit_forcing_window()
    run_steps(W, e2e=True)  # primes the pipeline: the first timed window's slices are in flight
    e2e_count.update(h2d=0, d2h=0)
    ms_e2e = timed(args.steps, e2e=True)
    h.synchronize()
    e2e_value = total_members * args.steps / (ms_e2e * 1e-3)
    h2d = e2e_count["h2d"] // args.steps
    d2h = e2e_count["d2h"] // args.steps

    # per-kernel durations (CUDA events on the library's stream) over the same number of steps
    # (one member group = one stream, so that kernels do not overlap and every event pair brackets
    # exactly one kernel)
    h.set_groups(1)
    sp0 = h.stats()
    h.profile(True)
    run_steps(args.steps)
    report = h.profile_report()
    h.profile(False)
    sp1 = h.stats()
    h.set_groups(args.groups)
    roof = roofline(flat, sym, report, sp0, sp1, nm, args)

    # final result gather (the only collective of the design)
    dev_buf = torch.empty((nb, nm), dtype=torch.float64, device="cuda")
    h.copy_member_to_device("storage", dev_buf.data_ptr(), nb)
    gathered = gather_members(dev_buf, total_members) if world > 1 else dev_buf.cpu().numpy()

    if rank == 0:
        assert gathered is not None and gathered.shape == (nb, total_members)
        member_steps = nm * args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, flat, sym, nm, world),
            "simulated_days_per_s": value * DT / 86400.0,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "steps_per_call": W,
                    "how": "rbs_stage_forcing_window (pinned host forcing slices of every outer step) + "
                           "rbs_advance_window (pinned host storages of every outer step out)"},
            "gpu_launches": l1 - l0,
            "clocks": clocks,
            "newton_iters_per_member_step": (s1["newton_iters"] - s0["newton_iters"]) / member_steps,
            "substeps_per_member_step": (s1["substeps"] - s0["substeps"]) / member_steps,
            "rejects_per_member_step": (s1["rejects"] - s0["rejects"]) / member_steps,
            "passes_per_step": (s1["passes"] - s0["passes"]) / args.steps,
            "negative_storage_member_steps": int(model.status_counts["negative_storage"]),
            "min_storage": float(gathered.min()),
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = args.cpu_members or min(nm, 256 * cores)
            cpu_steps = min(args.steps, 24)
            val, el, st = cpu_arm(p, flat, sample, cpu_steps, args.spinup + args.warmup,
                                  args.full_newton, cores)
            line["cpu_baseline"] = {
                "value": val, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{sample} members x {cpu_steps} steps after {args.spinup + args.warmup} "
                          f"spin-up steps; C++ oracle port, one instance per member on {cores} "
                          f"host threads ({el:.1f} s)"}
        print(json.dumps(line), flush=True)
    model.finalize()
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(group: str):
    """dram read+write bytes per launch of the group's main kernel from the committed
    `ncu --set full` capture (profiles/r1_*.txt; full 4096-member launch), or None."""
    main = {"linear_solve": "k_newton_linear_ent", "rhs_jacobian": "k_links",
            "newton_update": "k_newton_update", "limiter_accept": "k_step_apply"}.get(group)
    f = ROOT / "profiles" / f"r1_{main}.txt"
    if main is None or not f.exists():
        return None
    for line in f.read_text().splitlines():
        if line.startswith("# dram traffic"):
            val, unit = line.split(":")[1].split()[:2]
            return float(val) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    return None


def roofline(flat, sym, report, s0, s1, nm, args):
    """HBM roofline of the dominant kernel group: algorithmic bytes (SURVEY §8(d) per-member
    figures x member-launches actually executed) / summed CUDA-event duration."""
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6500.0, "fallback (B200_PROFILING.md)"
    alg = algorithmic_bytes(flat, sym)
    iters = s1["newton_iters"] - s0["newton_iters"]
    attempts = (s1["substeps"] - s0["substeps"]) + (s1["rejects"] - s0["rejects"])
    jac_evals = iters if args.full_newton else attempts

    def ms_of(*names):
        return sum(v[1] for k, v in report.items() if any(nm_ in k for nm_ in names))

    groups = {
        # RHS + analytic Jacobian (fused): basin, link and controller kernels
        "rhs_jacobian": (ms_of("k_basin<", "k_links", "k_pid", "k_continuous_control"),
                         alg["rhs"] * iters + alg["jac"] * jac_evals),
        # factorisation + substitution (one fused shared-memory kernel on the step path)
        "linear_solve": (ms_of("k_lu_", "k_solve_", "k_newton_linear", "k_assemble_linear"),
                         alg["solve"] * iters + (0 if args.full_newton else alg["lu_spill"] * iters)),
        "newton_update": (ms_of("k_newton_update", "k_newton_converge"),
                          (32 * flat["n_state"] + 8 * flat["nnz_j"] + 8 * flat["n_reduced"]) * iters),
        "limiter_accept": (ms_of("k_basin_limit", "k_step_", "k_compact"), alg["limiter"] * attempts),
    }
    total_ms = sum(v[1] for v in report.values())
    name, (gms, gbytes) = max(groups.items(), key=lambda kv: kv[1][0])
    achieved = gbytes / (gms * 1e-3) / 1e9 if gms > 0 else 0.0
    step_bytes = sum(b for _, b in groups.values())
    return {
        "bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": ncu_traffic(name), "peak_source": peak_src,
        "share_of_step": gms / total_ms if total_ms else None,
        "step_achieved": step_bytes / (total_ms * 1e-3) / 1e9 if total_ms else None,
        "groups_ms": {k: round(v[0], 3) for k, v in groups.items()},
        "groups_gbs": {k: round(v[1] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None
                       for k, v in groups.items()},
        "kernels_ms": {k: [v[0], round(v[1], 3)] for k, v in sorted(report.items())},
        "how": "same number of steps repeated on a single stream with one CUDA-event pair per "
               "launch; bytes = per-member algorithmic bytes x member-launches executed",
    }


if __name__ == "__main__":
    main()
