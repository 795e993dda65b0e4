/* ribasim_b200.h — C ABI of librbs.so, the B200 (sm_100a) implicit-integration inner loop
 * for Ribasim networks, batched over ensemble members.
 *
 * This is the drop-in boundary for the hot path only (level B3).  A Julia
 * host keeps the model, TOML/GeoPackage input, BMI and callbacks, and `ccall`s these entry
 * points from inside the SciML plug-in seams of the reference (INTEGRATION.md shows the
 * binding).  Each entry point cites the reference interface it replaces, relative to
 * /root/reference.
 *
 * Conventions
 *  - plain C linkage, plain pointers and sizes; no torch / C++ types cross the boundary;
 *  - every function returning `int` returns 0 on success, non-zero on failure; the message
 *    is fetched with rbs_last_error (mirrors libribasim's get_last_bmi_error,
 *    build/libribasim.jl:40-52, 192-201);
 *  - host buffers are borrowed for the duration of the call; the library owns all device
 *    memory; one handle per device; calls on one handle are serialised by the caller and are
 *    stream-ordered inside the library;
 *  - per-member arrays are structure-of-arrays `[entity][member]`, member fastest
 *    (`value[e * n_members + m]`), Float64; indices are 0-based Int32;
 *  - state order is the reference's `state_components` (core/src/parameter.jl:11-22);
 *    the reduced state is `[basins sorted by node_id ; PID integrals]`
 *    (core/src/read.jl:1758-1766).
 *  - there is no CPU fallback: rbs_create fails loudly when no CUDA device is usable.
 */
#ifndef RIBASIM_B200_H
#define RIBASIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rbs_builder rbs_builder; /* network description under construction */
typedef struct rbs_handle rbs_handle;   /* one ensemble shard resident on one device */

/* ---- network description ------------------------------------------------------------
 * Replaces the in-memory `ParametersIndependent` the Julia core builds in
 * `Parameters(db, config)` (core/src/read.jl:1692-1802).  The host fills named Int32 /
 * Float64 arrays (names listed in DESIGN.md §"descriptor keys"; produced by
 * ribasim_b200/flatten.py and ribasim_b200/symbolic.py) and hands the builder to rbs_create.
 * Optional tuning keys (one int32 each): "n_groups", "coop_members", "ent_tile", "lin_m",
 * "lin_lane" (warps per block of the member-per-lane solve, 0 = off, default 32) and
 * "lin_lane_min" (launches of at least this many members take it, default 2048), "dc_rec_cap". */
rbs_builder* rbs_builder_create(void);
int rbs_builder_set_i32(rbs_builder* b, const char* key, const int32_t* data, int64_t n);
int rbs_builder_set_f64(rbs_builder* b, const char* key, const double* data, int64_t n);
void rbs_builder_destroy(rbs_builder* b);

/* Allocate device state for `n_members` ensemble members on CUDA device `device`.
 * Returns NULL on failure (see rbs_last_error).  Replaces `Model(config)`'s allocation of
 * u, du, caches and the linear-solve cache (core/src/model.jl:174-237,
 * core/src/differentiation.jl:278-300). */
rbs_handle* rbs_create(const rbs_builder* b, int32_t n_members, int32_t device);
void rbs_destroy(rbs_handle* h);

/* Copy the last error message (NUL terminated) into buf; returns its length. */
int rbs_last_error(char* buf, int32_t len);
/* Library version string. */
const char* rbs_version(void);
/* Number of kernels launched on this handle since creation (for benchmarks). */
int64_t rbs_kernel_launches(const rbs_handle* h);

/* ---- parameters and forcing ---------------------------------------------------------
 * Shared (member-independent) parameter arrays, e.g. "lb_level", "pump_flow_rate",
 * "trc_table_idx": what `eval_time_interpolation` caches per RHS time in the reference
 * (core/src/util.jl:1216-1231, core/src/parameter.jl:812-842), and the values swapped by
 * DiscreteControl / allocation (core/src/callback.jl:765-786). */
int rbs_set_param_f64(rbs_handle* h, const char* key, const double* host, int64_t n);
int rbs_set_param_i32(rbs_handle* h, const char* key, const int32_t* host, int64_t n);
/* Parameter ensembles: a per-member override `[n_entity][n_members]` of a shared link parameter
 * ("mn_manning_n", "lr_resistance", "pump_flow_rate", "outlet_flow_rate", "lb_level": the fields
 * of ManningResistance / LinearResistance / Pump / Outlet / LevelBoundary,
 * core/src/parameter.jl:513-772, that one perturbs in a calibration, uncertainty or boundary-
 * scenario ensemble).  A node under DiscreteControl keeps reading
 * its control-state table.  host = NULL drops the override (all members use the shared value
 * again).  A member whose override equals the shared value gets bitwise the shared-value result. */
int rbs_set_member_param_f64(rbs_handle* h, const char* key, const double* host, int64_t n_entity);
/* Per-member arrays `[n_entity][n_members]`: Basin vertical fluxes ("precipitation",
 * "surface_runoff", "potential_evaporation", "drainage", "infiltration" — the BMI-writable
 * forcing arrays, core/src/bmi.jl:64-88), "fb_rate", state "u", exact cumulatives.
 * `broadcast != 0`: host holds one value per entity that is replicated to every member. */
int rbs_set_member_f64(rbs_handle* h, const char* key, const double* host, int64_t n_entity,
                       int32_t broadcast);
int rbs_get_member_f64(rbs_handle* h, const char* key, double* host, int64_t n_entity);
/* Asynchronous variants for PINNED host buffers, to overlap transfers with stepping.
 * rbs_set_member_f64_async stages the upload on a copy stream into a shadow buffer while the
 * next rbs_step_implicit_euler / rbs_advance call computes with the current values; the staged
 * values become live when that call returns, i.e. they are the input of the step after it
 * (stage step k+1's forcing, then run step k).  The host buffer may be reused after that call.
 * rbs_get_member_f64_async snapshots the array on the device (ordered after the work enqueued so
 * far) and transfers it in the background; the host buffer is valid after rbs_synchronize. */
int rbs_set_member_f64_async(rbs_handle* h, const char* key, const double* pinned_host, int64_t n_entity);
int rbs_get_member_f64_async(rbs_handle* h, const char* key, double* pinned_host, int64_t n_entity);
/* Device-to-device copy of a per-member array into caller-owned device memory (e.g. the send
 * buffer of the final NCCL gather of an ensemble sharded over GPUs). */
int rbs_copy_member_to_device(rbs_handle* h, const char* key, void* dst_device, int64_t n_entity);
/* Integer per-member arrays: "status", "dc_state" (control state index per DiscreteControl
 * node, -1 = none yet), "dc_truth" (per condition), counters "cnt_iters", "cnt_sub", "cnt_rej",
 * "cnt_dc" (control state changes).  The DiscreteControl record of the reference
 * (`discrete_control.record`, core/src/callback.jl:698-716) is kept per member for the first 64
 * changes: entry r < cnt_dc[m] is ("dc_rec_t" f64 [64][n_members] model time, "dc_rec_ns" i32 =
 * DiscreteControl index | control state index << 16, "dc_rec_bits" i32 = truth state bits). */
int rbs_get_member_i32(rbs_handle* h, const char* key, int32_t* host, int64_t n_entity);
/* DiscreteControl at the initial time (`apply_discrete_control!` with t == 0,
 * core/src/callback.jl:608-661): evaluates every condition on the caches of the last
 * rbs_refresh and sets the control state of every member.  During stepping the evaluation runs on
 * the device after every accepted (sub-)step of a member; per-member control states select the
 * row of the control-state parameter tables ("pump_cs_flow_rate", ...). */
int rbs_apply_discrete_control(rbs_handle* h);
/* Previous accepted time `tprev` (core/src/parameter.jl:1126-1130). */
int rbs_set_tprev(rbs_handle* h, double tprev);

/* ---- SciML plug-in seams (host buffers in, host buffers out) ------------------------ */
/* u_reduced = A*u : `reduce_state!` (core/src/util.jl:745-791).  u: [n_state][B]. */
int rbs_reduce(rbs_handle* h, const double* u, double* u_reduced);
/* du = g(u_reduced, t): `water_balance!` (core/src/solve.jl:35-84).  Optional outputs
 * (may be NULL): the basin caches current_storage/level/area/low_storage_factor
 * [n_basin][B] (core/src/parameter.jl:793-803). */
int rbs_rhs(rbs_handle* h, const double* u_reduced, double t, double* du, double* storage,
            double* level, double* area, double* factor);
/* Analytic J_intermediate = dg/du_reduced in the fixed CSR pattern (values [nnz_j][B]) and
 * W_inner = A*J_intermediate - I/gamma (values [nnz_w][B]); replaces `get_jacobian!`
 * (core/src/differentiation.jl:361-390) and `calc_W_inner!` (:153-271).  Either output may
 * be NULL.  The factorisation of W_inner is refreshed as a side effect. */
int rbs_jac(rbs_handle* h, const double* u_reduced, double t, double gamma, double* j_vals,
            double* w_vals);
/* Solve W_inner x = b with the current factors; b, x: [n_reduced][B].  Replaces the KLU
 * solve inside `dolinsolve` (core/src/differentiation.jl:307-359). */
int rbs_solve(rbs_handle* h, const double* b, double* x);
/* `limit_flow!` (core/src/solve.jl:840-984): clamp u given uprev, dt at time t. */
int rbs_limit_flow(rbs_handle* h, double* u, const double* uprev, double dt, double t);

/* ---- device-resident time stepping ---------------------------------------------------
 * Advance every member from t to t + dt by implicit Euler with NLNewton on the state resident
 * in the handle ("u"): what OrdinaryDiffEq's perform_step!/nlsolve! do through the seams above
 * (core/src/model.jl:219-237), followed by limit_flow!, the negative-storage
 * check (isoutofdomain, core/src/util.jl:918-924) and the exact-forcing bookkeeping of
 * `update_cumulative_flows!` (core/src/callback.jl:125-182).
 * Each member runs its own sub-step sequence inside [t, t + dt]: an attempt whose Newton
 * iteration fails or that ends with negative storage is rejected and retried with half the
 * length, an accepted attempt that converged quickly doubles the nominal length again (what
 * the reference's adaptive controller does on a failed nonlinear solve).  Members never
 * influence each other, so results do not depend on how the ensemble is sharded.
 * status_per_member (may be NULL): 0 ok, |1 a forced accept after Newton failure at the
 * minimum sub-step, |2 negative storage accepted at the minimum sub-step. */
int rbs_step_implicit_euler(rbs_handle* h, double t, double dt, int32_t* status_per_member);
/* The same for a window of n_steps consecutive outer steps of length dt during which forcing
 * and shared parameters do not change (e.g. the hourly steps between two daily forcing updates,
 * the solver's tstops: core/src/model.jl:166,194-196).  Every member walks through its own
 * n_steps outer steps without waiting for the others at the step boundaries; the per-member
 * result is identical to n_steps calls of rbs_step_implicit_euler.  status bits are OR-ed over
 * the window. */
int rbs_advance(rbs_handle* h, double t, double dt, int32_t n_steps, int32_t* status_per_member);
/* Forced window: n_steps outer steps whose forcing differs per step, with the per-step results
 * streamed back.  This is the batched form of the reference's main loop, where the forcing of the
 * whole horizon is known up front (`update_basin!` at the forcing tstops, core/src/callback.jl:
 * 846-866, and the `saveat` storage output, core/src/callback.jl:125-182): every member walks
 * through the window at its own pace reading the forcing slice of the outer step it is in, so
 * members that need many sub-steps do not stall the ensemble at every step boundary.
 * rbs_stage_forcing_window uploads, asynchronously from PINNED host memory, the slices
 * [n_steps][n_entity][n_members] of outer steps first_step .. first_step + n_steps - 1 of the
 * window (a window may be staged in several chunks) of one forcing array ("precipitation", "surface_runoff",
 * "potential_evaporation", "drainage", "infiltration", "fb_rate") into a shadow buffer; the
 * staged slices become live when the next rbs_advance_window returns (so: stage window k+1, then
 * run window k, and the upload overlaps with the computation) or at once with
 * rbs_commit_forcing_window (before the first window).  A live set of slices serves one
 * rbs_advance_window; arrays without live slices keep their current value for the whole window.
 * rbs_advance_window runs the
 * window and, when storage_out_pinned is not NULL, starts the background download of the basin
 * storages at the end of every outer step, [n_steps][n_basin][n_members]; with flow_out_pinned
 * likewise the mean flow over every outer step of every state (increment of the cumulative
 * state / dt: what `save_flow` stores per saveat interval, core/src/callback.jl:390-456),
 * [n_steps][n_state][n_members].  The slices of an outer step go out as soon as every member has
 * completed it, while the stragglers of the window are still computing; the buffers are valid
 * after rbs_synchronize (or after the next but one rbs_advance_window returns).  Results are identical to rbs_set_member_f64 +
 * rbs_step_implicit_euler per step. */
int rbs_stage_forcing_window(rbs_handle* h, const char* key, const double* pinned_host,
                             int64_t n_entity, int32_t first_step, int32_t n_steps);
int rbs_commit_forcing_window(rbs_handle* h);
/* Forcing that changes less often than the outer step (the reference's forcing tstops are the
 * times of the Basin / time table, core/src/callback.jl:846-866 — daily rows under an hourly
 * dt): outer step k of a window reads slice (k + phase) / steps_per_slice, so the `first_step` /
 * `n_steps` of rbs_stage_forcing_window count slices and a day is uploaded once, not 24 times.
 * Default 1, 0: one slice per outer step.  Applies to the windows advanced after the call. */
int rbs_set_forcing_slicing(rbs_handle* h, int32_t steps_per_slice, int32_t phase);
int rbs_advance_window(rbs_handle* h, double t, double dt, int32_t n_steps,
                       double* storage_out_pinned, double* flow_out_pinned,
                       int32_t* status_per_member);
/* Solver options: abstol, reltol (core/src/config.jl:187-188), max Newton iterations. */
int rbs_set_solver(rbs_handle* h, double abstol, double reltol, int32_t max_iter);
/* Stepper options.  full_newton: 0 = J and W evaluated once per attempt at (uprev, t + h) like
 * the reference's nlsolve!, 1 = refreshed every Newton iteration.  predictor: 0 = z0 = 0
 * (OrdinaryDiffEq `extrapolant = :constant`), 1 = last accepted increment rescaled to the new
 * sub-step.  max_halvings: smallest sub-step is dt * 2^-max_halvings (0 = plain fixed-step
 * implicit Euler).  grow_iters: double the sub-step after converging within this many
 * iterations.  early_exit: give up an attempt as soon as the contraction rate theta predicts
 * that eta*|dz| cannot fall below kappa within max_iter (Hairer-Wanner) instead of iterating to
 * max_iter.  shrink_log2: a rejected attempt is retried with h / 2^shrink_log2.
 * Defaults: 1, 1, 20, 4, 1, 2 (with the analytic Jacobian a refresh costs about as much as an
 * RHS evaluation, and full Newton needs fewer iterations and far fewer rejected sub-steps than
 * the frozen variant on stiff networks). */
int rbs_set_stepper(rbs_handle* h, int32_t full_newton, int32_t predictor, int32_t max_halvings,
                    int32_t grow_iters, int32_t early_exit, int32_t shrink_log2);
/* Number of member groups (0 = automatic).  Each group runs its own state machine on its own
 * CUDA stream so that the latency-bound kernels of one group overlap with the others'; results
 * do not depend on the grouping. */
int rbs_set_groups(rbs_handle* h, int32_t n_groups);
/* Counters since creation: [0] outer steps, [1] kernel passes, [2] accepted sub-steps summed
 * over members, [3] member-steps flagged |1, [4] Newton iterations summed over members,
 * [5] rejected attempts summed over members. */
int rbs_get_stats(rbs_handle* h, int64_t* out6);
/* Refresh the basin caches (storage, level, area, factor) and du at the resident state. */
int rbs_refresh(rbs_handle* h, double t);

/* Per-launch timing with CUDA events on the handle's stream: rbs_profile(h, 1) starts
 * recording every kernel launch, rbs_profile_report writes one line per kernel
 * "name<TAB>launches<TAB>total_ms" (NUL terminated, truncated to len), rbs_profile(h, 0) stops. */
int rbs_profile(rbs_handle* h, int32_t enable);
int rbs_profile_report(rbs_handle* h, char* buf, int32_t len);

/* Tuning aid: average duration (ms) of the Newton linear-solve kernel (W assembly from the
 * current J values, factorisation, substitution) on the first n_active members, `reps` back to
 * back launches on the handle's stream.  Leaves the stepping state machine reset. */
int rbs_time_linear(rbs_handle* h, int32_t n_active, int32_t reps, double* ms_per_launch);

/* Page-locked host memory for the asynchronous and forced-window entry points (a host language
 * without its own pinned allocator, e.g. Julia via unsafe_wrap, takes its staging buffers from
 * here).  NULL on failure (rbs_last_error tells why). */
void* rbs_host_alloc(int64_t bytes);
void rbs_host_free(void* p);

/* Raw device pointer of a per-member array (for zero-copy interop, e.g. NCCL gathers or
 * torch tensors wrapping library memory); NULL if unknown. */
void* rbs_device_ptr(rbs_handle* h, const char* key);
/* The CUDA stream all work of this handle is ordered on (cudaStream_t as void*). */
void* rbs_stream(rbs_handle* h);
/* For an integrator that lives above the seams (adaptive QNDF, core/src/config.jl:368-380): an
 * accepted step of length dt advances the exactly integrated forcing — cumulative precipitation,
 * surface runoff, drainage and boundary inflow (`update_cumulative_flows!`, core/src/callback.jl:
 * 131-170) — with the rates in force, and tprev by dt; rbs_rhs / rbs_jac at a later time then see
 * S = S0 + A u + the exact inflows up to that time (core/src/solve.jl:137-150). */
int rbs_accept_step(rbs_handle* h, double dt);
/* Final result gather, the only collective of the design (members are sharded over
 * the GPUs of a box and never exchange anything inside the time loop; the reference has no
 * analogue, one `Ribasim.Model` is one process): ncclAllGather of the member array `key`
 * ("storage", "u", ...) over the ranks of the caller's communicator (`ncclComm_t`, created by the
 * host with NCCL.jl / ncclCommInitRank) on the handle's stream.  Every rank must hold the same
 * number of members; recv_device receives [n_ranks][n_entity][n_members_local] doubles (rank
 * blocks in rank order, each in the library's member-fastest layout).  NCCL is bound at call time. */
int rbs_gather(rbs_handle* h, void* nccl_comm, const char* key, int64_t n_entity, void* recv_device);
int rbs_synchronize(rbs_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* RIBASIM_B200_H */
