"""ctypes binding of librbs.so (include/ribasim_b200.h) and its build recipe.

The product path has no CPU fallback: `load()` raises if the CUDA library is missing and
`Handle.create` raises if no device is usable.
"""

from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np

from .flatten import Flat
from .symbolic import Symbolic, analyse, build_entries

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
# RBS_LIB selects another build of the same sources (e.g. the -DRBS_LIN_PROFILE tuning build)
LIB_PATH = Path(os.environ["RBS_LIB"]) if os.environ.get("RBS_LIB") else PKG_DIR / "librbs.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # keep the reference's operation order (no FMA contraction)
    "-Xcompiler", "-fPIC", "-shared",
]


class RbsError(RuntimeError):
    pass


def _sources() -> list[Path]:
    return [CSRC / "rbs.cu", CSRC / "rbs_kernels.cuh", CSRC / "rbs_coop.cuh", CSRC / "rbs_tile.cuh",
            CSRC / "rbs_fused.cuh", CSRC / "rbs_lane.cuh", CSRC / "rbs_math.cuh", PKG_DIR.parent / "include" / "ribasim_b200.h"]


def _source_hash() -> str:
    import hashlib

    hsh = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for s in _sources():
        hsh.update(s.read_bytes())
    return hsh.hexdigest()


def _hash_file() -> Path:
    return LIB_PATH.with_name(LIB_PATH.name + ".srchash")


def is_stale() -> bool:
    """True when librbs.so was not built from the sources in the tree (content hash of the sources
    and flags stored next to the library; robust against copies that do not keep mtimes)."""
    hf = _hash_file()
    return not LIB_PATH.exists() or not hf.exists() or hf.read_text().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile librbs.so for sm_100a in-tree (so that it travels with the repo snapshot)."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC / "rbs.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RbsError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    _hash_file().write_text(_source_hash() + "\n")
    if verbose:
        print(res.stderr)
    return LIB_PATH


SHIM_PATH = PKG_DIR / "libribasim.so"


def build_shim(force: bool = False) -> Path:
    """Compile the `libribasim` stand-in (csrc/libribasim_shim.c, include/libribasim.h): the C ABI
    of /root/reference/build/libribasim.jl:74-212 over the embedded-Python BMI of this package."""
    import sysconfig

    src = CSRC / "libribasim_shim.c"
    if not force and SHIM_PATH.exists() and SHIM_PATH.stat().st_mtime >= src.stat().st_mtime:
        return SHIM_PATH
    inc = sysconfig.get_config_var("INCLUDEPY")
    libdir = sysconfig.get_config_var("LIBDIR")
    ver = sysconfig.get_config_var("LDVERSION")
    site = sysconfig.get_paths()["purelib"]
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-fvisibility=hidden", f"-I{inc}",
           f'-DRBS_REPO_ROOT="{PKG_DIR.parent}"', f'-DRBS_SITE_PACKAGES="{site}"',
           "-o", str(SHIM_PATH), str(src), f"-L{libdir}", f"-lpython{ver}", "-ldl",
           f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RbsError(f"gcc failed:\n{res.stdout}\n{res.stderr}")
    return SHIM_PATH


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("RBS_LIB") or os.environ.get("RBS_NO_BUILD"):
        if not LIB_PATH.exists():
            raise RbsError(f"{LIB_PATH} is missing (run __graft_entry__.build())")
    else:
        build()  # no-op unless the sources changed since the library was built
    L = ctypes.CDLL(str(LIB_PATH))
    vp, cp = ctypes.c_void_p, ctypes.c_char_p
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    i64, i32, f64 = ctypes.c_int64, ctypes.c_int32, ctypes.c_double
    sig = {
        "rbs_builder_create": (vp, []),
        "rbs_builder_set_i32": (ctypes.c_int, [vp, cp, ip, i64]),
        "rbs_builder_set_f64": (ctypes.c_int, [vp, cp, dp, i64]),
        "rbs_builder_destroy": (None, [vp]),
        "rbs_create": (vp, [vp, i32, i32]),
        "rbs_destroy": (None, [vp]),
        "rbs_last_error": (ctypes.c_int, [cp, i32]),
        "rbs_version": (cp, []),
        "rbs_kernel_launches": (i64, [vp]),
        "rbs_set_param_f64": (ctypes.c_int, [vp, cp, dp, i64]),
        "rbs_set_param_i32": (ctypes.c_int, [vp, cp, ip, i64]),
        "rbs_set_member_param_f64": (ctypes.c_int, [vp, cp, dp, i64]),
        "rbs_set_member_f64": (ctypes.c_int, [vp, cp, dp, i64, i32]),
        "rbs_get_member_f64": (ctypes.c_int, [vp, cp, dp, i64]),
        "rbs_set_member_f64_async": (ctypes.c_int, [vp, cp, dp, i64]),
        "rbs_get_member_f64_async": (ctypes.c_int, [vp, cp, dp, i64]),
        "rbs_stage_forcing_window": (ctypes.c_int, [vp, cp, dp, i64, i32, i32]),
        "rbs_advance_window": (ctypes.c_int, [vp, f64, f64, i32, dp, dp, ip]),
        "rbs_commit_forcing_window": (ctypes.c_int, [vp]),
        "rbs_set_forcing_slicing": (ctypes.c_int, [vp, i32, i32]),
        "rbs_host_alloc": (vp, [i64]),
        "rbs_host_free": (None, [vp]),
        "rbs_copy_member_to_device": (ctypes.c_int, [vp, cp, vp, i64]),
        "rbs_get_member_i32": (ctypes.c_int, [vp, cp, ip, i64]),
        "rbs_apply_discrete_control": (ctypes.c_int, [vp]),
        "rbs_set_tprev": (ctypes.c_int, [vp, f64]),
        "rbs_reduce": (ctypes.c_int, [vp, dp, dp]),
        "rbs_rhs": (ctypes.c_int, [vp, dp, f64, dp, dp, dp, dp, dp]),
        "rbs_jac": (ctypes.c_int, [vp, dp, f64, f64, dp, dp]),
        "rbs_solve": (ctypes.c_int, [vp, dp, dp]),
        "rbs_limit_flow": (ctypes.c_int, [vp, dp, dp, f64, f64]),
        "rbs_step_implicit_euler": (ctypes.c_int, [vp, f64, f64, ip]),
        "rbs_advance": (ctypes.c_int, [vp, f64, f64, i32, ip]),
        "rbs_set_solver": (ctypes.c_int, [vp, f64, f64, i32]),
        "rbs_set_stepper": (ctypes.c_int, [vp, i32, i32, i32, i32, i32, i32]),
        "rbs_set_groups": (ctypes.c_int, [vp, i32]),
        "rbs_get_stats": (ctypes.c_int, [vp, ctypes.POINTER(i64)]),
        "rbs_refresh": (ctypes.c_int, [vp, f64]),
        "rbs_time_linear": (ctypes.c_int, [vp, i32, i32, dp]),
        "rbs_profile": (ctypes.c_int, [vp, i32]),
        "rbs_profile_report": (ctypes.c_int, [vp, cp, i32]),
        "rbs_device_ptr": (vp, [vp, cp]),
        "rbs_stream": (vp, [vp]),
        "rbs_synchronize": (ctypes.c_int, [vp]),
        "rbs_gather": (ctypes.c_int, [vp, vp, cp, i64, vp]),
        "rbs_accept_step": (ctypes.c_int, [vp, f64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # raises AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "rbs_builder_create", "rbs_builder_set_i32", "rbs_builder_set_f64", "rbs_builder_destroy",
    "rbs_create", "rbs_destroy", "rbs_last_error", "rbs_version", "rbs_kernel_launches",
    "rbs_set_param_f64", "rbs_set_param_i32", "rbs_set_member_param_f64", "rbs_set_member_f64", "rbs_get_member_f64", "rbs_set_member_f64_async",
    "rbs_get_member_f64_async", "rbs_stage_forcing_window", "rbs_advance_window", "rbs_commit_forcing_window", "rbs_set_forcing_slicing", "rbs_host_alloc", "rbs_host_free",
    "rbs_copy_member_to_device", "rbs_get_member_i32", "rbs_apply_discrete_control", "rbs_set_tprev", "rbs_reduce", "rbs_rhs", "rbs_jac", "rbs_solve", "rbs_limit_flow",
    "rbs_step_implicit_euler", "rbs_advance", "rbs_set_solver", "rbs_set_stepper", "rbs_set_groups", "rbs_get_stats", "rbs_refresh",
    "rbs_time_linear", "rbs_profile", "rbs_profile_report", "rbs_device_ptr", "rbs_stream", "rbs_synchronize", "rbs_gather", "rbs_accept_step",
]


def _dp(a: np.ndarray | None):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a: np.ndarray | None):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def last_error() -> str:
    L = load()
    buf = ctypes.create_string_buffer(1025)
    L.rbs_last_error(buf, 1025)
    return buf.value.decode(errors="replace")


def _check(rc: int) -> None:
    if rc != 0:
        raise RbsError(last_error())


# parameters refreshed by the host whenever the RHS time crosses a tstop
PARAM_F64 = [
    "lb_level", "trc_max_downstream_level",
    "pump_flow_rate", "pump_min_flow_rate", "pump_max_flow_rate", "pump_min_upstream_level",
    "pump_max_downstream_level", "pump_flow_demand",
    "outlet_flow_rate", "outlet_min_flow_rate", "outlet_max_flow_rate",
    "outlet_min_upstream_level", "outlet_max_downstream_level", "outlet_flow_demand",
    "ud_total_demand", "ud_return_factor", "pid_target", "pid_proportional", "pid_integral",
    "pid_derivative", "cc_sv_const", "dc_thr_hi", "dc_thr_lo", "dc_sv_const",
]
PARAM_I32 = ["trc_table_idx"]


class PinnedArray:
    """A float64 numpy array over page-locked host memory from `rbs_host_alloc` (staging buffer of
    the asynchronous / forced-window calls for hosts without torch)."""

    def __init__(self, shape):
        L = load()
        self.shape = tuple(int(x) for x in shape)
        n = int(np.prod(self.shape))
        self._L = L
        self.ptr = L.rbs_host_alloc(max(n, 1) * 8)
        if not self.ptr:
            raise RbsError(last_error())
        buf = (ctypes.c_double * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(self.shape)

    def close(self) -> None:
        if getattr(self, "ptr", None):
            self.array = None
            self._L.rbs_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Handle:
    """One ensemble shard on one GPU (`rbs_handle`)."""

    def __init__(self, flat: Flat, sym: Symbolic | None = None, n_members: int = 1,
                 device: int = 0, lu_mode: int = -1, lu_tile: int = 0, smem_tile: int = -1,
                 smem_threads: int = 1024, lin_mode: int = -1, n_groups: int = 0,
                 coop_members: int = -1, tile_window: int = -1, lin_m: int = 0,
                 r2_pass: int = 0, fin_cluster: int = 4, dc_rec_cap: int = 64):
        self.L = load()
        self.flat = flat
        if sym is None:
            sym = analyse(flat.ints["n_reduced"], flat.arrays["wrow_ptr"], flat.arrays["wcol"])
        self.sym = sym
        self.B = int(n_members)
        self.n_state = flat.ints["n_state"]
        self.n_reduced = flat.ints["n_reduced"]
        self.n_basin = flat.ints["n_basin"]
        b = self.L.rbs_builder_create()
        try:
            def set_i(key, arr):
                a = np.ascontiguousarray(np.asarray(arr, dtype=np.int32)).reshape(-1)
                _check(self.L.rbs_builder_set_i32(b, key.encode(), _ip(a), a.size))

            def set_d(key, arr):
                a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64)).reshape(-1)
                _check(self.L.rbs_builder_set_f64(b, key.encode(), _dp(a), a.size))

            for k, v in flat.ints.items():
                set_i(k, [v])
            for k, v in flat.floats.items():
                set_d(k, [v])
            for k, v in flat.arrays.items():
                if v.dtype == np.int32:
                    set_i(k, v)
                else:
                    set_d(k, v)
            wdiag_flag = np.zeros(flat.ints["nnz_w"], dtype=np.int32)
            wdiag_flag[flat.arrays["wdiag"]] = 1
            set_i("wdiag_flag", wdiag_flag)
            set_i("n_lu", [sym.n_lu])
            for name in ("perm", "lu_wsrc", "lu_pivot", "lu_prod_ptr", "lu_prod_l", "lu_prod_u",
                         "lu_level_ptr", "lu_diag", "fwd_level_ptr", "fwd_row", "fwd_ptr", "fwd_l",
                         "fwd_k", "bwd_level_ptr", "bwd_row", "bwd_ptr", "bwd_u", "bwd_k",
                         "ldu_wsrc", "ldu_flag", "ldu_prod_ptr", "ldu_prod_l", "ldu_prod_u",
                         "ldu_prod_p", "ldu_level_ptr", "ldu_diag", "ldu_fwd_level_ptr",
                         "ldu_fwd_row", "ldu_fwd_ptr", "ldu_fwd_l", "ldu_fwd_k",
                         "ldu_bwd_level_ptr", "ldu_bwd_row", "ldu_bwd_ptr", "ldu_bwd_u",
                         "ldu_bwd_k"):
                set_i(name, getattr(sym, name))
            # entry schedule of the tile kernels (slots + schedule must fit the 227 KB of shared
            # memory of an sm_100 SM; the library checks and falls back to the other tiers)
            ent = build_entries(sym)
            if ent is not None:
                set_i("ent_e", ent.ent_e)
                set_i("ent_q0", ent.ent_q0)
                set_i("ent_cf", ent.ent_cf)
                set_i("ent_prod_l", ent.prod_l)
                set_i("ent_prod_p", ent.prod_p)
                set_i("ent_prod_u", ent.prod_u)
                set_i("ent_lvl_ptr", ent.lvl_ptr)
                set_i("ent_team_log", ent.team_log)
                set_i("ent_n_elim", [ent.n_elim])
                set_i("ent_n_bwd", [ent.n_bwd])
                set_i("ent_tile", [int(os.environ.get("RBS_ENT_TILE", "0"))])
                # the same schedule on compacted slots (fill-ins take over dead slots) for the
                # full-Newton path: fewer slots per tile, more tiles per SM
                entc = build_entries(sym, compact=True)
                if entc is not None and entc.n_slots < ent.n_slots:
                    set_i("entc_e", entc.ent_e)
                    set_i("entc_q0", entc.ent_q0)
                    set_i("entc_cf", entc.ent_cf)
                    set_i("entc_prod_l", entc.prod_l)
                    set_i("entc_prod_p", entc.prod_p)
                    set_i("entc_prod_u", entc.prod_u)
                    set_i("entc_lvl_ptr", entc.lvl_ptr)
                    set_i("entc_team_log", entc.team_log)
                    set_i("entc_load_e", entc.load_e)
                    set_i("entc_load_s", entc.load_s)
                    set_i("entc_n_slots", [entc.n_slots])
                    set_i("entc_y0", [entc.y0])
            set_i("lu_mode", [lu_mode])
            set_i("lu_tile", [lu_tile])
            set_i("smem_tile", [smem_tile])
            set_i("smem_threads", [smem_threads])
            set_i("lin_mode", [lin_mode])
            set_i("n_groups", [n_groups])
            # members per thread of the full-Newton solve (rbs_fused.cuh): 8 / 4, 0 = the 2-member
            # tiles of rbs_kernels.cuh (faster on B200, measured), -1 = the widest that fits
            set_i("lin_m", [lin_m])
            # 1: convergence test + finalize half as one cluster kernel per pass (rbs_fused.cuh,
            # `fin_cluster` blocks per 32-member tile); 0: the round-1 launch sequence
            set_i("r2_pass", [r2_pass])
            set_i("fin_cluster", [fin_cluster])
            # DiscreteControl state changes recorded per member (`discrete_control.record`)
            set_i("dc_rec_cap", [dc_rec_cap])
            # ensembles of at most this many members run a window as one cooperative kernel
            # (-1 = the library's default)
            set_i("coop_members", [int(os.environ.get("RBS_COOP_MEMBERS", str(coop_members)))])
            # > 0: a block per tile of that many members (2 or 4) runs a whole window
            # (rbs_tile.cuh); 0 = the stream path; -1 = the library's default
            set_i("tile_window", [int(os.environ.get("RBS_TILE_WINDOW", str(tile_window)))])
            self.h = self.L.rbs_create(b, self.B, int(device))
        finally:
            self.L.rbs_builder_destroy(b)
        if not self.h:
            raise RbsError(last_error())

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.rbs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def set_param(self, key: str, arr) -> None:
        if key in PARAM_I32 or (isinstance(arr, np.ndarray) and arr.dtype == np.int32):
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.int32)).reshape(-1)
            _check(self.L.rbs_set_param_i32(self.h, key.encode(), _ip(a), a.size))
        else:
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64)).reshape(-1)
            _check(self.L.rbs_set_param_f64(self.h, key.encode(), _dp(a), a.size))

    def set_time_params(self, tp: dict[str, np.ndarray]) -> None:
        for k in PARAM_F64:
            if k in tp:
                self.set_param(k, tp[k])
        for k in PARAM_I32:
            if k in tp:
                self.set_param(k, np.asarray(tp[k], dtype=np.int32))

    def set_member_param(self, key: str, arr) -> None:
        """Per-member override `[n_entity][B]` of a shared link parameter (None drops it)."""
        if arr is None:
            _check(self.L.rbs_set_member_param_f64(self.h, key.encode(), None, 0))
            return
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64))
        _check(self.L.rbs_set_member_param_f64(self.h, key.encode(), _dp(a.reshape(-1)), a.size // self.B))

    def set_member(self, key: str, arr, broadcast: bool = False) -> None:
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64))
        n_entity = a.size if broadcast else a.size // self.B
        if not broadcast and a.size != n_entity * self.B:
            raise ValueError(f"{key}: size {a.size} is not a multiple of n_members={self.B}")
        _check(self.L.rbs_set_member_f64(self.h, key.encode(), _dp(a.reshape(-1)), n_entity,
                                         int(broadcast)))

    def get_member(self, key: str, n_entity: int) -> np.ndarray:
        out = np.empty((n_entity, self.B), dtype=np.float64)
        _check(self.L.rbs_get_member_f64(self.h, key.encode(), _dp(out.reshape(-1)), n_entity))
        return out

    def set_member_ptr(self, key: str, host_ptr: int, n_entity: int) -> None:
        """Like set_member but from a raw host pointer (e.g. a pinned staging buffer)."""
        _check(self.L.rbs_set_member_f64(self.h, key.encode(),
                                         ctypes.cast(host_ptr, ctypes.POINTER(ctypes.c_double)),
                                         n_entity, 0))

    def get_member_ptr(self, key: str, host_ptr: int, n_entity: int) -> None:
        _check(self.L.rbs_get_member_f64(self.h, key.encode(),
                                         ctypes.cast(host_ptr, ctypes.POINTER(ctypes.c_double)),
                                         n_entity))

    def set_member_ptr_async(self, key: str, pinned_ptr: int, n_entity: int) -> None:
        """Stage an upload from pinned memory; live at the start of the next step."""
        _check(self.L.rbs_set_member_f64_async(
            self.h, key.encode(), ctypes.cast(pinned_ptr, ctypes.POINTER(ctypes.c_double)), n_entity))

    def get_member_ptr_async(self, key: str, pinned_ptr: int, n_entity: int) -> None:
        """Background download into pinned memory; valid after synchronize()."""
        _check(self.L.rbs_get_member_f64_async(
            self.h, key.encode(), ctypes.cast(pinned_ptr, ctypes.POINTER(ctypes.c_double)), n_entity))

    def stage_forcing_window(self, key: str, pinned_ptr: int, n_entity: int, n_steps: int,
                             first_step: int = 0) -> None:
        """Upload the forcing slices [n_steps][n_entity][B] of outer steps first_step.. of the next
        forced window (pinned host memory, asynchronous)."""
        _check(self.L.rbs_stage_forcing_window(
            self.h, key.encode(), ctypes.cast(pinned_ptr, ctypes.POINTER(ctypes.c_double)), n_entity,
            int(first_step), int(n_steps)))

    def set_forcing_slicing(self, steps_per_slice: int = 1, phase: int = 0) -> None:
        """Outer step k of a forced window reads slice (k + phase) // steps_per_slice (daily forcing
        under hourly steps: 24, step-of-day of the window's first step)."""
        _check(self.L.rbs_set_forcing_slicing(self.h, int(steps_per_slice), int(phase)))

    def commit_forcing_window(self) -> None:
        """Make the staged slices live now (before the first window)."""
        _check(self.L.rbs_commit_forcing_window(self.h))

    def advance_window(self, t: float, dt: float, n_steps: int, storage_out_ptr: int | None = None,
                       want_status: bool = True, flow_out_ptr: int | None = None) -> np.ndarray | None:
        """Forced window: staged per-step forcing in; per-step storages and mean flows out
        (pinned host buffers, filled in the background)."""
        st = np.zeros(self.B, dtype=np.int32) if want_status else None
        dpt = ctypes.POINTER(ctypes.c_double)
        out = ctypes.cast(storage_out_ptr, dpt) if storage_out_ptr else None
        fout = ctypes.cast(flow_out_ptr, dpt) if flow_out_ptr else None
        _check(self.L.rbs_advance_window(self.h, float(t), float(dt), int(n_steps), out, fout,
                                         _ip(st) if want_status else None))
        return st

    def copy_member_to_device(self, key: str, device_ptr: int, n_entity: int) -> None:
        _check(self.L.rbs_copy_member_to_device(self.h, key.encode(), ctypes.c_void_p(device_ptr),
                                                n_entity))

    def stream_ptr(self) -> int:
        return int(self.L.rbs_stream(self.h) or 0)

    def get_member_i32(self, key: str, n_entity: int) -> np.ndarray:
        out = np.empty((n_entity, self.B), dtype=np.int32)
        _check(self.L.rbs_get_member_i32(self.h, key.encode(), _ip(out.reshape(-1)), n_entity))
        return out

    def apply_discrete_control(self) -> None:
        _check(self.L.rbs_apply_discrete_control(self.h))

    def set_tprev(self, t: float) -> None:
        _check(self.L.rbs_set_tprev(self.h, float(t)))

    def set_solver(self, abstol: float, reltol: float, max_iter: int = 10) -> None:
        _check(self.L.rbs_set_solver(self.h, abstol, reltol, max_iter))

    # ------------------------------------------------------------------ seams
    def _mat(self, x, n_entity: int) -> np.ndarray:
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        if a.size != n_entity * self.B:
            raise ValueError(f"expected {n_entity} x {self.B} values, got {a.size}")
        return a.reshape(n_entity, self.B)

    def reduce(self, u) -> np.ndarray:
        u = self._mat(u, self.n_state)
        out = np.empty((self.n_reduced, self.B))
        _check(self.L.rbs_reduce(self.h, _dp(u), _dp(out)))
        return out

    def rhs(self, u_reduced, t: float, with_cache: bool = False):
        ur = self._mat(u_reduced, self.n_reduced)
        du = np.empty((self.n_state, self.B))
        if with_cache:
            c = {k: np.empty((self.n_basin, self.B)) for k in ("storage", "level", "area", "factor")}
            _check(self.L.rbs_rhs(self.h, _dp(ur), t, _dp(du), _dp(c["storage"]), _dp(c["level"]),
                                  _dp(c["area"]), _dp(c["factor"])))
            return du, c
        _check(self.L.rbs_rhs(self.h, _dp(ur), t, _dp(du), None, None, None, None))
        return du

    def jac(self, u_reduced, t: float, gamma: float):
        ur = self._mat(u_reduced, self.n_reduced)
        jv = np.empty((self.flat.ints["nnz_j"], self.B))
        wv = np.empty((self.flat.ints["nnz_w"], self.B))
        _check(self.L.rbs_jac(self.h, _dp(ur), t, gamma, _dp(jv), _dp(wv)))
        return jv, wv

    def solve(self, b) -> np.ndarray:
        b = self._mat(b, self.n_reduced)
        x = np.empty_like(b)
        _check(self.L.rbs_solve(self.h, _dp(b), _dp(x)))
        return x

    def limit_flow(self, u, uprev, dt: float, t: float) -> np.ndarray:
        u = self._mat(u, self.n_state).copy()
        up = self._mat(uprev, self.n_state)
        _check(self.L.rbs_limit_flow(self.h, _dp(u), _dp(up), dt, t))
        return u

    def step(self, t: float, dt: float, want_status: bool = True):
        st = np.zeros(self.B, dtype=np.int32) if want_status else None
        _check(self.L.rbs_step_implicit_euler(self.h, t, dt, _ip(st)))
        return st

    def advance(self, t: float, dt: float, n_steps: int, want_status: bool = True):
        """A window of `n_steps` outer steps with constant forcing (rbs_advance)."""
        st = np.zeros(self.B, dtype=np.int32) if want_status else None
        _check(self.L.rbs_advance(self.h, t, dt, int(n_steps), _ip(st)))
        return st

    def accept_step(self, dt: float) -> None:
        """Advance the exactly integrated forcing and tprev by an accepted step (rbs_accept_step)."""
        _check(self.L.rbs_accept_step(self.h, float(dt)))

    def refresh(self, t: float) -> None:
        _check(self.L.rbs_refresh(self.h, t))

    def set_stepper(self, full_newton: bool = True, predictor: bool = True, max_halvings: int = 20,
                    grow_iters: int = 4, early_exit: bool = True, shrink_log2: int = 2) -> None:
        _check(self.L.rbs_set_stepper(self.h, int(full_newton), int(predictor), int(max_halvings),
                                      int(grow_iters), int(early_exit), int(shrink_log2)))

    def set_groups(self, n_groups: int) -> None:
        _check(self.L.rbs_set_groups(self.h, int(n_groups)))

    def stats(self) -> dict[str, int]:
        out = (ctypes.c_int64 * 6)()
        _check(self.L.rbs_get_stats(self.h, out))
        return dict(steps=out[0], passes=out[1], substeps=out[2], failed=out[3],
                    newton_iters=out[4], rejects=out[5])

    def time_linear(self, n_active: int, reps: int = 10) -> float:
        out = ctypes.c_double(0.0)
        _check(self.L.rbs_time_linear(self.h, int(n_active), int(reps), ctypes.byref(out)))
        return out.value

    def profile(self, enable: bool) -> None:
        _check(self.L.rbs_profile(self.h, int(enable)))

    def profile_report(self) -> dict[str, tuple[int, float]]:
        """{kernel name: (launches, total milliseconds)} since profiling was enabled."""
        buf = ctypes.create_string_buffer(1 << 16)
        _check(self.L.rbs_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split("\t")
            out[name] = (int(cnt), float(ms))
        return out

    def kernel_launches(self) -> int:
        return int(self.L.rbs_kernel_launches(self.h))

    def device_ptr(self, key: str) -> int:
        return int(self.L.rbs_device_ptr(self.h, key.encode()) or 0)

    def gather(self, nccl_comm: int, key: str, n_entity: int, recv_device_ptr: int) -> None:
        """ncclAllGather of a member array over the caller's communicator (rbs_gather)."""
        _check(self.L.rbs_gather(self.h, ctypes.c_void_p(nccl_comm), key.encode(), int(n_entity),
                                 ctypes.c_void_p(recv_device_ptr)))

    def synchronize(self) -> None:
        _check(self.L.rbs_synchronize(self.h))
