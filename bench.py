#!/usr/bin/env python
"""Benchmark of the implicit-integration inner loop on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, host cores)

A "step" is one outer implicit-Euler step (dt = 3600 s) of every ensemble member: all Newton
iterations, sub-steps, limit_flow!, the negative-storage check and the exact-forcing bookkeeping.

Headline workload = BASELINE.json config 4: the HWS-like synthetic network (the national HWS
model is a private DVC artefact, SURVEY intro fact 5; `models/hws/ribasim.toml` is used instead
when it exists) with an ensemble of 4096 perturbed-forcing members IN TOTAL, sharded over the N
GPUs (strong scaling: 512 members per GPU at N = 8).  Members are independent, so ranks share
nothing inside the loop and the only collective is the final gather.

Rank 0 prints ONE JSON line.  Under `extra_configs` the same line carries BASELINE config 2
(backwater, single instance) and config 5 (random tree of 100k basins, 256 members in total), each
with its own roofline and CPU baseline; under `extra.weak` (N > 1) the weak-scaling figure with
4096 members per GPU.
"""

from __future__ import annotations

import argparse
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
STEPS_PER_DAY = 24
METRIC = "scenario-steps/sec"
UNIT = "scenario-steps/s"
FORCING_KEYS = ("precipitation", "potential_evaporation", "drainage", "fb_rate")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96, help="timed outer steps (default: 4 simulated days)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=0,
                    help="ensemble members IN TOTAL (0 = the workload's BASELINE size: 4096 hws, 256 tree, 1 backwater)")
    ap.add_argument("--weak", action="store_true", help="--members is per GPU (weak scaling)")
    ap.add_argument("--basins", type=int, default=0, help="0 = the workload's BASELINE size")
    ap.add_argument("--workload", default="hws", choices=["hws", "tree", "backwater"],
                    help="hws: HWS-like network (BASELINE config 4); tree: random tree (config 5); "
                         "backwater: Manning chain, single instance (config 2)")
    ap.add_argument("--spinup", type=int, default=-1, help="untimed steps from the initial state (-1 = per workload)")
    ap.add_argument("--full-newton", type=int, default=1)
    ap.add_argument("--lu-tile", type=int, default=0)
    ap.add_argument("--lu-mode", type=int, default=-1)
    ap.add_argument("--smem-tile", type=int, default=-1)
    ap.add_argument("--smem-threads", type=int, default=1024)
    ap.add_argument("--lin-mode", type=int, default=-1)
    ap.add_argument("--groups", type=int, default=0, help="member groups on own streams (0 = auto)")
    ap.add_argument("--e2e-window", type=int, default=96,
                    help="outer steps per rbs_advance_window call of the end-to-end leg")
    ap.add_argument("--cpu-members", type=int, default=0, help="sample size of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel timing pass")
    ap.add_argument("--ncu-range", action="store_true",
                    help="bracket the device-resident timed leg with cudaProfilerStart/Stop (for "
                         "`ncu --profile-from-start off`) and stop after it")
    return ap.parse_args()


WORKLOAD_DEFAULTS = {
    # workload: (basins, members in total, spin-up steps)
    "hws": (500, 4096, 240),
    "tree": (100_000, 256, 120),
    "backwater": (51, 1, 240),
}


# ----------------------------------------------------------------------------- workload
def build_workload(n_basins: int, workload: str = "hws"):
    from ribasim_b200.flatten import flatten
    from ribasim_b200.network import build_params
    from ribasim_b200.symbolic import analyse
    from ribasim_b200.tables import ModelTables
    from ribasim_b200.testmodels import MODELS, hws_like_model, random_tree_model

    source = "synthetic"
    if workload == "tree":
        tables = random_tree_model(n_basins=n_basins, n_hours=24 * 30)
    elif workload == "backwater":
        tables = MODELS["backwater"]()
    else:
        real = ROOT / "models" / "hws" / "ribasim.toml"  # the real model if present
        if real.exists():
            tables, source = ModelTables.read(real), "models/hws"
        else:
            tables = hws_like_model(n_basins=n_basins, n_days=30)
    p = build_params(tables)
    flat = flatten(p)
    sym = analyse(flat["n_reduced"], flat["wrow_ptr"], flat["wcol"])
    return p, flat, sym, source


def base_forcing(p, day: int):
    """Unperturbed basin vertical fluxes and boundary inflows of calendar day `day`."""
    t = day * 86400.0
    # a missing value keeps the flux it replaces (0 at the start: core/src/callback.jl:826-838)
    base = {k: np.nan_to_num(np.array([s(t) for s in p.basin.forcing[k]], dtype=float)) for k in
            ("precipitation", "potential_evaporation", "drainage")}
    fb = np.array([s(t) for s in p.nodes["flow_boundary"]["flow_rate"]], dtype=float)
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
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons))


# ----------------------------------------------------------------------------- CPU arm
def ctypes_double(x: float):
    import ctypes

    return ctypes.c_double(x)


def cpu_arm(p, flat, n_members: int, n_steps: int, spinup: int, threads: int, perturb: bool = True,
            warm: dict | None = None):
    """Oracle port (C++ restatement of the reference path) on host threads: one instance per
    member, the reference's own Newton variant (J and W evaluated once per attempt,
    OrdinaryDiffEq NLNewton) — also the faster one on the CPU.  `warm`: start the members from a
    state computed elsewhere (the device's state after its spin-up, where a CPU spin-up would take
    minutes): {"step": outer steps done, "u", "cum_precipitation", "cum_surface_runoff",
    "cum_drainage", "fb_cum": [n_entity][n_members], "h": [n_members], "eta": [n_members]}."""
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
    batch = OracleBatch(members, advance_opts(1e-5, 1e-5, 10, 0, 4, 20, 1, 1, 2), threads)

    def set_day(day):
        base, fb = base_forcing(p, day)
        if perturb:
            f = perturbed_forcing(base, fb, range(n_members), day)
        else:
            f = {k: np.repeat(v[:, None], n_members, axis=1) for k, v in base.items()}
            f["fb_rate"] = np.repeat(fb[:, None], n_members, axis=1)
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
            if step % STEPS_PER_DAY == 0:
                set_day(step // STEPS_PER_DAY)
            chunk = min(k - done, STEPS_PER_DAY - step % STEPS_PER_DAY)
            s = batch.run(t, DT, chunk)
            for key in stats:
                stats[key] += s[key]
            t += chunk * DT
            step += chunk
            done += chunk
        return stats

    if warm is not None:
        step = int(warm["step"])
        t = step * DT
        for j, om in enumerate(members):
            om.set_f64("cumulative_precipitation", np.ascontiguousarray(warm["cum_precipitation"][:, j]))
            om.set_f64("cumulative_surface_runoff", np.ascontiguousarray(warm["cum_surface_runoff"][:, j]))
            om.set_f64("cumulative_drainage", np.ascontiguousarray(warm["cum_drainage"][:, j]))
            if warm["fb_cum"].shape[0]:
                om.set_f64("fb_cumulative_flow", np.ascontiguousarray(warm["fb_cum"][:, j]))
            om.L.orc_set_tprev(om.h, ctypes_double(t))
            batch.us[j][:] = warm["u"][:, j]
        batch.h_try[:] = warm["h"]
        batch.eta[:] = warm["eta"]
        set_day(step // STEPS_PER_DAY)
        spinup = 0
    elif spinup == 0:
        set_day(0)
    advance_steps(spinup)
    t0 = time.perf_counter()
    stats = advance_steps(n_steps)
    el = time.perf_counter() - t0
    return n_members * n_steps / el, el, stats


def resolve(args, workload: str):
    """(basins, members in total, spin-up) of a workload after the command-line overrides."""
    nb, nm, sp = WORKLOAD_DEFAULTS[workload]
    if workload == args.workload:
        nb = args.basins or nb
        nm = args.members or nm
        sp = args.spinup if args.spinup >= 0 else sp
    return nb, nm, sp


def cpu_sample(workload: str, cores: int, requested: int, members_local: int) -> tuple[int, int]:
    """Bounded sample (members, steps) of the CPU arm: ~10-30 s of host work."""
    if workload == "tree":
        return requested or min(cores, 16), 2
    if workload == "backwater":
        return 1, 240
    # ~10-30 s of host work at ~1.7 k member-steps/s per core, spin-up included
    return requested or min(members_local, 64 * cores), 96


def reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm for this path on the box's host cores.
    Julia is not installed (SURVEY intro fact 4), so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nb, nm, spinup = resolve(args, args.workload)
    p, flat, sym, source = build_workload(nb, args.workload)
    sample, steps = cpu_sample(args.workload, cores, args.cpu_members, nm)
    steps = min(steps, args.steps)
    val, el, stats = cpu_arm(p, flat, sample, steps, spinup + args.warmup, cores,
                             perturb=args.workload != "backwater")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / steps,
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64", "data": source,
        "config": workload_config(args.workload, flat, sym, nm, nm, 1, spinup, args.full_newton, source)
        | {"sample_members": sample, "sample_steps": steps},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} members x {steps} steps after "
                                   f"{spinup + args.warmup} spin-up steps, one oracle "
                                   f"instance per member on {cores} host threads ({el:.1f} s)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "newton_iters_per_member_step": stats["iters"] / (sample * steps),
        "substeps_per_member_step": stats["substeps"] / (sample * steps),
    }
    print(json.dumps(line), flush=True)


def workload_config(workload, flat, sym, members_local, members_total, world, spinup, full_newton, source):
    names = {"hws": "hws_like synthetic network" if source == "synthetic" else "models/hws",
             "tree": "random-tree synthetic network", "backwater": "backwater Manning chain (ribasim_testmodels)"}
    return {
        "workload": f"{names[workload]}, {flat['n_basin']} basins x {members_total} members in total "
                    f"({members_local} on this GPU), dt=3600 s implicit Euler",
        "n_basin": flat["n_basin"], "n_state": flat["n_state"], "n_reduced": flat["n_reduced"],
        "nnz_j": flat["nnz_j"], "nnz_w": flat["nnz_w"], "n_lu": sym.n_lu,
        "members_per_gpu": members_local, "members_total": members_total,
        "dt_s": DT, "spinup_steps": spinup, "full_newton": full_newton,
        "l2": "per-step working set (~0.2 MB of per-member state x members) exceeds the 126 MB L2 from "
              "~600 members per GPU; below that the kernels are latency bound, not cache bound",
        "parallelism": f"members sharded over {world} GPU(s), no collective in the loop",
    }


# ----------------------------------------------------------------------------- GPU arm
class Run:
    """One workload on this rank's GPU: spin-up, device-resident timing, end-to-end timing,
    per-kernel timing."""

    def __init__(self, args, workload, lo, hi, total_members, local_rank, world, dist):
        import torch
        from ribasim_b200.model import Model

        self.torch, self.dist, self.world = torch, dist, world
        self.args, self.workload = args, workload
        nb_req, _, self.spinup = resolve(args, workload)
        self.p, self.flat, self.sym, self.source = build_workload(nb_req, workload)
        self.lo, self.hi, self.nm, self.total = lo, hi, hi - lo, total_members
        self.perturb = workload != "backwater"
        flat = self.flat
        self.nb, self.n, self.nfb = flat["n_basin"], flat["n_state"], flat["n_fb"]
        self.model = Model(self.p, n_members=self.nm, device=local_rank, dt=DT, flat=flat, sym=self.sym,
                           full_newton=bool(args.full_newton), lu_tile=args.lu_tile, lu_mode=args.lu_mode,
                           smem_tile=args.smem_tile, smem_threads=args.smem_threads,
                           lin_mode=args.lin_mode, n_groups=args.groups)
        self.h = self.model.h
        self.stream = torch.cuda.ExternalStream(self.h.stream_ptr(), device=local_rank)
        self.stage: dict[int, dict] = {}  # day -> pinned per-member forcing arrays
        self.t, self.step = 0.0, 0
        self.warm = None
        self.model._overridden.update(("precipitation", "potential_evaporation", "drainage"))
        self.h.set_forcing_slicing(STEPS_PER_DAY, 0)

    def pinned(self, shape):
        return self.torch.empty(shape, dtype=self.torch.float64).pin_memory()

    def fill_day(self, day):
        """Host-side generation of one day of (perturbed) forcing into pinned staging buffers (the
        ensemble's input producer; kept outside every timed region)."""
        from ribasim_b200.ensemble import perturbed_forcing

        if day in self.stage:
            return
        self.stage[day] = {k: self.pinned((self.nfb if k == "fb_rate" else self.nb, self.nm)) for k in FORCING_KEYS}
        base, fb = base_forcing(self.p, day)
        out = {k: v.numpy() for k, v in self.stage[day].items()}
        if self.perturb:
            perturbed_forcing(base, fb, range(self.lo, self.hi), day, out=out)
        else:
            for k in ("precipitation", "potential_evaporation", "drainage"):
                out[k][:] = base[k][:, None]
            out["fb_rate"][:] = fb[:, None]

    def stage_days(self, step0, n_steps):
        """H2D of the daily forcing slices that outer steps step0 .. step0 + n_steps - 1 read (one
        slice per day and array: rbs_set_forcing_slicing); returns the bytes uploaded."""
        d0, d1 = step0 // STEPS_PER_DAY, (step0 + n_steps - 1) // STEPS_PER_DAY
        nbytes = 0
        for d in range(d0, d1 + 1):
            self.fill_day(d)
        for key in FORCING_KEYS:
            for d in range(d0, d1 + 1):
                src = self.stage[d][key]
                if src.shape[0] == 0:
                    continue
                self.h.stage_forcing_window(key, src.data_ptr(), src.shape[0], 1, first_step=d - d0)
                nbytes += src.shape[0] * self.nm * 8
        return nbytes

    def window(self, n_steps, out_ptr=None):
        """n_steps outer steps as ONE forced window on the live slices."""
        self.h.set_forcing_slicing(STEPS_PER_DAY, self.step % STEPS_PER_DAY)
        self.h.advance_window(self.t, DT, n_steps, out_ptr, want_status=False)
        self.t += n_steps * DT
        self.step += n_steps

    def resident(self, n_steps):
        """Device-resident leg: forcing of the whole window staged beforehand, results stay in HBM."""
        self.stage_days(self.step, n_steps)
        self.h.commit_forcing_window()
        self.h.synchronize()
        return lambda: self.window(n_steps)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, sync_copies=False):
        """fn() bracketed by barrier + synchronize on both sides, CUDA events on the library's
        stream, max over ranks (ms)."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record(self.stream)
        fn()
        if sync_copies:
            self.h.synchronize()  # the download of the last window's results belongs to the timed region
        ev1.record(self.stream)
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        if self.world > 1:
            tt = torch.tensor([ms], device="cuda")
            self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    def spin_up(self, warmup):
        """Untimed: from the initial state through `spinup` steps (windows of a day), then the
        warm-up steps."""
        left = self.spinup
        while left > 0:
            w = min(left, STEPS_PER_DAY - self.step % STEPS_PER_DAY)
            self.resident(w)()
            left -= w
        if warmup > 0:
            self.resident(warmup)()
        self.h.synchronize()

    def measure(self, steps, warmup, local_rank, profile=True):
        args, h = self.args, self.h
        self.spin_up(warmup)
        self.warm = self.snapshot(16) if self.workload == "tree" and self.world == 1 else None
        # -- device-resident: inputs in HBM when the timed region starts
        run = self.resident(steps)
        sampler = ClockSampler(local_rank)
        sampler.start()
        s0, l0 = h.stats(), h.kernel_launches()
        if args.ncu_range:
            self.torch.cuda.profiler.start()
        ms = self.timed(run)
        if args.ncu_range:
            self.torch.cuda.profiler.stop()
        s1, l1 = h.stats(), h.kernel_launches()
        value = self.total * steps / (ms * 1e-3)
        if args.ncu_range:
            sampler.stop()
            print(f"profiled leg: {steps} steps, {ms / steps:.4f} ms/step, {l1 - l0} launches, "
                  f"{(s1['newton_iters'] - s0['newton_iters']) / (self.nm * steps):.3f} Newton iterations per "
                  f"member-step", file=sys.stderr)
            raise SystemExit(0)
        # -- end to end through the C ABI with HOST buffers: every window uploads its daily forcing
        # slices from pinned memory and downloads the storages of every outer step; the upload of
        # window k+1 and the download of window k-1 overlap with window k
        W = max(1, min(args.e2e_window if self.workload != "tree" else 6, steps))
        results = [self.pinned((W, self.nb, self.nm)) for _ in range(2)]
        counts = {"h2d": 0, "d2h": 0, "calls": 0}

        def e2e_steps(k):
            done = 0
            while done < k:
                w = min(W, k - done)
                nxt = min(W, k - done - w) or W
                counts["h2d"] += self.stage_days(self.step + w, nxt)  # live when this window returns
                self.window(w, results[counts["calls"] % 2].data_ptr())
                counts["d2h"] += w * self.nb * self.nm * 8
                counts["calls"] += 1
                done += w

        for d in [d for d in self.stage if d < self.step // STEPS_PER_DAY]:
            del self.stage[d]
        for d in range(self.step // STEPS_PER_DAY, (self.step + steps + 3 * W) // STEPS_PER_DAY + 2):
            self.fill_day(d)
        self.stage_days(self.step, W)
        h.commit_forcing_window()
        e2e_steps(W)  # primes the pipeline: the first timed window's slices are in flight
        counts.update(h2d=0, d2h=0)
        ms_e2e = self.timed(lambda: e2e_steps(steps), sync_copies=True)
        h.synchronize()
        clocks = sampler.stop()
        e2e_value = self.total * steps / (ms_e2e * 1e-3)
        out = {
            "value": value, "ms": ms, "steps": steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": counts["h2d"] // steps,
                    "d2h_bytes_per_step": counts["d2h"] // steps, "ms_per_step": ms_e2e / steps,
                    "steps_per_call": W,
                    "how": "rbs_stage_forcing_window (pinned host forcing, one slice per simulated day and "
                           "array) + rbs_advance_window (pinned host storages of every outer step out)"},
            "gpu_launches": l1 - l0, "clocks": clocks, "stats": (s0, s1),
        }
        # -- per-kernel durations (CUDA events on the library's stream) over one simulated day (one
        # member group = one stream, so that kernels do not overlap and every event pair brackets
        # exactly one kernel)
        if profile:
            psteps = min(steps, STEPS_PER_DAY)
            h.set_groups(1)
            run = self.resident(psteps)
            sp0 = h.stats()
            h.profile(True)
            run()
            report = h.profile_report()
            h.profile(False)
            sp1 = h.stats()
            h.set_groups(args.groups)
            out["roofline"] = roofline(self.flat, self.sym, report, sp0, sp1, args)
        return out

    def snapshot(self, n_members: int) -> dict:
        """State of the first members at the current time, for a warm start of the CPU arm."""
        h, k = self.h, min(n_members, self.nm)
        f = self.flat
        get = lambda key, ne: h.get_member(key, ne)[:, :k].copy()
        return {"step": self.step, "u": get("u", f["n_state"]),
                "cum_precipitation": get("cum_precipitation", f["n_basin"]),
                "cum_surface_runoff": get("cum_surface_runoff", f["n_basin"]),
                "cum_drainage": get("cum_drainage", f["n_basin"]),
                "fb_cum": get("fb_cum", f["n_fb"]) if f["n_fb"] else np.zeros((0, k)),
                "h": h.get_member("mhnom", 1)[0, :k].copy(), "eta": h.get_member("eta", 1)[0, :k].copy()}

    def gather(self):
        """Final result gather (the only collective of the design)."""
        from ribasim_b200.ensemble import gather_members

        torch = self.torch
        dev_buf = torch.empty((self.nb, self.nm), dtype=torch.float64, device="cuda")
        self.h.copy_member_to_device("storage", dev_buf.data_ptr(), self.nb)
        return gather_members(dev_buf, self.total) if self.world > 1 else dev_buf.cpu().numpy()

    def close(self):
        self.model.finalize()


def result_fields(run: Run, m: dict, world: int, weak: bool) -> dict:
    s0, s1 = m["stats"]
    member_steps = run.nm * m["steps"]
    return {
        "value": m["value"], "unit": UNIT, "ms_per_step": m["ms"] / m["steps"], "steps": m["steps"],
        "scaling": "weak" if weak else "strong",
        "config": workload_config(run.workload, run.flat, run.sym, run.nm, run.total, world, run.spinup,
                                  run.args.full_newton, run.source),
        "simulated_days_per_s": m["value"] * DT / 86400.0,
        "e2e": m["e2e"], "gpu_launches": m["gpu_launches"],
        "newton_iters_per_member_step": (s1["newton_iters"] - s0["newton_iters"]) / member_steps,
        "substeps_per_member_step": (s1["substeps"] - s0["substeps"]) / member_steps,
        "rejects_per_member_step": (s1["rejects"] - s0["rejects"]) / member_steps,
        "passes_per_step": (s1["passes"] - s0["passes"]) / m["steps"],
        "negative_storage_member_steps": int(run.model.status_counts["negative_storage"]),
        **({"roofline": m["roofline"]} if "roofline" in m else {}),
    }


def cpu_baseline_for(run: Run, args, warm: dict | None = None) -> dict:
    cores = os.cpu_count() or 1
    sample, steps = cpu_sample(run.workload, cores, args.cpu_members if run.workload == args.workload else 0, run.nm)
    if warm is not None:
        sample = min(sample, warm["u"].shape[1])
        warm = {k: (v[..., :sample] if isinstance(v, np.ndarray) else v) for k, v in warm.items()}
    val, el, _ = cpu_arm(run.p, run.flat, sample, steps, run.spinup + args.warmup, cores, perturb=run.perturb,
                         warm=warm)
    start = "from the device's state after its spin-up" if warm is not None else "after its own spin-up"
    return {"value": val, "unit": UNIT, "cores": min(cores, sample), "kind": "port",
            "sample": f"{sample} members x {steps} steps {start}; C++ oracle port, one instance per "
                      f"member on {min(cores, sample)} host threads ({el:.1f} s)"}


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

    from ribasim_b200.ensemble import member_range

    def shard(total, weak):
        if weak:
            return rank * total, (rank + 1) * total, total * world
        lo, hi = member_range(total, rank, world)
        return lo, hi, total

    # ---- headline workload
    _, nm_total, _ = resolve(args, args.workload)
    lo, hi, total = shard(nm_total, args.weak)
    if hi - lo < 1:
        raise SystemExit("fewer members than GPUs")
    run = Run(args, args.workload, lo, hi, total, local_rank, world, dist)
    steps = args.steps if args.workload != "tree" else min(args.steps, 24)
    m = run.measure(steps, args.warmup, local_rank, profile=not args.no_profile)
    gathered = run.gather()
    line = None
    if rank == 0:
        assert gathered is not None and gathered.shape == (run.nb, total)
        f = result_fields(run, m, world, args.weak)
        line = {
            "metric": METRIC, "value": f.pop("value"), "unit": f.pop("unit"), "n_gpus": world,
            "steps": f.pop("steps"), "warmup": args.warmup, "ms_per_step": f.pop("ms_per_step"),
            "higher_is_better": True, "scaling": f.pop("scaling"), "vs_baseline": None, "dtype": "f64",
            "data": run.source, **f, "clocks": m["clocks"], "min_storage": float(gathered.min()),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_for(run, args, run.warm)
    run.close()
    del run

    # ---- weak-scaling figure next to the strong-scaling headline
    extra = {}
    if world > 1 and not args.weak and not args.no_extra and args.workload == "hws":
        lo, hi, total = shard(nm_total, True)
        r2 = Run(args, "hws", lo, hi, total, local_rank, world, dist)
        m2 = r2.measure(min(args.steps, 48), args.warmup, local_rank, profile=False)
        if rank == 0:
            f = result_fields(r2, m2, world, True)
            extra["weak"] = {k: f[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "e2e",
                                               "passes_per_step")} | {"members_per_gpu": r2.nm, "members_total": total}
        r2.close()
        del r2

    # ---- the other BASELINE configs, each to the same measurement bar
    extra_configs = {}
    if not args.no_extra and args.workload == "hws":
        for wl in ("backwater", "tree"):
            _, nm_wl, _ = resolve(args, wl)
            if wl == "backwater":
                if rank != 0:
                    continue  # a single instance does not shard: "replicas only"
                lo, hi, total = 0, 1, 1
                wworld = 1
            else:
                lo, hi, total = shard(nm_wl, False)
                wworld = world
                if hi - lo < 1:
                    continue
            try:
                saved_world = world
                rx = Run(args, wl, lo, hi, total, local_rank, wworld, dist)
                mx = rx.measure(min(args.steps, 24) if wl == "tree" else 240, args.warmup, local_rank,
                                profile=not args.no_profile)
                if rank == 0:
                    f = result_fields(rx, mx, wworld, False)
                    f["metric"], f["n_gpus"], f["clocks"] = METRIC, wworld, mx["clocks"]
                    if wl == "backwater":
                        f["single_instance_ms_per_step"] = f["ms_per_step"]
                    if world == 1 and not args.no_cpu_baseline:
                        f["cpu_baseline"] = cpu_baseline_for(rx, args, rx.warm)
                    extra_configs[wl] = f
                rx.close()
                del rx, saved_world
            except Exception as e:  # an extra config must not take the headline line down
                if rank == 0:
                    extra_configs[wl] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        if extra:
            line["extra"] = extra
        if extra_configs:
            line["extra_configs"] = extra_configs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic(group: str):
    """Summed dram read+write bytes per launch of the kernels of the group from the committed
    `ncu --set full` captures of full 4096-member launches (profiles/r2_*.txt, else r1_*), or None."""
    lane = (ROOT / "profiles" / "r2_k_newton_linear_lane.txt").exists()  # full launches take the lane kernel
    kernels = {"linear_solve": ["k_newton_linear_lane" if lane else "k_newton_linear_ent", "k_assemble_linear"],
               "rhs_jacobian": ["k_basin", "k_links"],
               "newton_update": ["k_newton_update"],
               # first finalize half of the profiled window: 2001 of the 4096 members (decide / compaction: negligible)
               "limiter_accept": ["k_basin_limit_pre", "k_step_limit", "k_basin_limit_post", "k_step_apply"]}.get(group)
    if not kernels:
        return None
    total = 0.0
    for kname in kernels:
        f = next((ROOT / "profiles" / f"{r}_{kname}.txt" for r in ("r2", "r1")
                  if (ROOT / "profiles" / f"{r}_{kname}.txt").exists()), None)
        if f is None:
            return None
        for ln in f.read_text().splitlines():
            if ln.startswith("# dram traffic"):
                val, unit = ln.split(":")[1].split()[:2]
                total += float(val) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
                break
        else:
            return None
    return total


def roofline(flat, sym, report, s0, s1, args):
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
        "limiter_accept": (ms_of("k_basin_limit", "k_step_", "k_compact", "k_finalize", "k_discrete_control"),
                           alg["limiter"] * attempts),
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
        "how": "one simulated day repeated on a single stream with one CUDA-event pair per launch; "
               "bytes = per-member algorithmic bytes x member-launches executed; traffic = summed ncu "
               "DRAM bytes of the group's kernels for one launch each out of profiles/r2_k_*.txt (Newton half: full "
               "4096-member launches; limiter / accept: the first finalize half of the profiled window, 2001 members)",
    }


if __name__ == "__main__":
    main()
