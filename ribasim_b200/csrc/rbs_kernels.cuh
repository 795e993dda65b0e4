// rbs_kernels.cuh — sm_100a kernels of the batched implicit-integration inner loop.
//
// Layout: every per-member quantity is structure-of-arrays `[entity][member]`, member
// fastest, so a warp reads 32 consecutive members of one entity: fully coalesced 256 B
// requests, network topology / link parameters / profile tables are warp-uniform broadcast
// loads served from L1.  With a single member the same index arithmetic degenerates to
// `[entity]`, still coalesced over entities.
//
// Determinism: every output element is written by exactly one thread and every sum runs in a
// fixed order taken from the host-built index lists.  No floating-point atomics anywhere.
#pragma once
#include <stdint.h>

#include "rbs_math.cuh"

enum { KIND_BASIN = 0, KIND_LEVEL_BOUNDARY = 1, KIND_TERMINAL = 2, KIND_NONE = 3 };
enum { CT_TRC = 0, CT_PUMP = 1, CT_OUTLET = 2, CT_UD_IN = 3, CT_UD_OUT = 4, CT_LR = 5, CT_MANNING = 6 };
enum { CONTROL_NONE = 0, CONTROL_CONTINUOUS = 1, CONTROL_PID = 2 };
enum { PHASE_NEWTON = 0, PHASE_LIMIT = 1, PHASE_DONE = 2 };
enum { ACTION_NONE = 0, ACTION_ACCEPT = 1, ACTION_REJECT = 2 };

struct Net {
    // sizes
    int B;  // members on this device
    // launch domain: the first NA entries of `alist` (nullptr = identity); set per launch
    int NA;
    const int* alist;
    int n_basin, n_pid, n_state, n_conn, n_reduced;
    int n_trc, n_pump, n_outlet, n_ud, n_ud_link, n_lr, n_manning, n_lb, n_fb, n_cc, n_tab;
    int off_trc, off_pump, off_outlet, off_ud_in, off_ud_out, off_lr, off_manning, off_evap,
        off_infil, off_integral;
    int nnz_j, nnz_w, n_lu;
    int n_dc, n_dc_cond;
    double ldt;  // level_difference_threshold
    // basins (shared)
    const double *storage0, *low_storage_threshold, *fixed_area;
    const int* prof_ptr;
    const double *prof_level, *prof_dSdh, *prof_dSdh_slope, *prof_cumS, *prof_area, *prof_area_slope;
    const int *inc_ptr, *inc_state;
    const double* inc_sign;
    const int *bfb_ptr, *bfb_idx;
    const int* basin_obs;  // 1 = storage / area of this basin are read by a controller during the RHS
    // parameter ensembles: per-member overrides [entity][member] of shared link parameters
    // (nullptr = every member uses the shared value); set with rbs_set_member_param_f64
    const double *pm_mn_manning_n, *pm_lr_resistance, *pm_pump_flow_rate, *pm_outlet_flow_rate,
        *pm_lb_level;
    // connectors (shared)
    const int *conn_type, *conn_idx, *src_kind, *src_idx, *dst_kind, *dst_idx;
    const double *lr_resistance, *lr_max_flow_rate;
    const double *mn_length, *mn_manning_n, *mn_profile_width, *mn_profile_slope,
        *mn_upstream_bottom, *mn_downstream_bottom;
    const int* tab_ptr;
    const double *tab_t, *tab_u, *tab_du, *tab_c1, *tab_c2, *tab_slope_left, *tab_slope_right;
    const int *pump_control_type, *outlet_control_type;
    const int *ud_link_ptr, *ud_of_link;
    const double *ud_min_level, *ud_link_alloc;
    // time-dependent shared parameters (host refreshed)
    const double* lb_level;
    const int* trc_table_idx;
    const double* trc_max_downstream_level;
    const double *pump_flow_rate, *pump_min_flow_rate, *pump_max_flow_rate,
        *pump_min_upstream_level, *pump_max_downstream_level, *pump_flow_demand;
    const double *outlet_flow_rate, *outlet_min_flow_rate, *outlet_max_flow_rate,
        *outlet_min_upstream_level, *outlet_max_downstream_level, *outlet_flow_demand;
    const double *ud_total_demand, *ud_return_factor;
    const double *pid_target, *pid_proportional, *pid_integral, *pid_derivative;
    const double* cc_sv_const;
    // control structure (shared)
    const int *pid_listen_basin, *pid_target_is_outlet, *pid_target_idx, *pid_target_state;
    const int *pid_jpos_listen, *pid_jpos_integral, *pid_int_jpos;
    const int *pid_term_ptr, *pid_term_state;
    const double* pid_term_sign;
    const int *pid_jterm_ptr, *pid_jterm_src, *pid_jterm_dst;
    const double* pid_jterm_sign;
    const int *cc_target_is_outlet, *cc_target_idx, *cc_target_state, *cc_sv_ptr, *cc_sv_kind,
        *cc_sv_idx;
    const double* cc_sv_weight;
    const int *cc_term_ptr, *cc_term_kind, *cc_term_sv, *cc_term_src, *cc_term_dst;
    int cc_tab_offset;
    // Jacobian / W patterns (shared)
    const int *jrow_ptr, *jcol, *jpos_src, *jpos_dst, *ud_out_jpos;
    const int *wsrc_ptr, *wsrc_pos, *wdiag_flag;
    const double* wsrc_sign;
    // LU schedule (shared)
    const int *lu_wsrc, *lu_pivot, *lu_prod_ptr, *lu_prod_l, *lu_prod_u, *lu_diag, *perm;
    const int *fwd_row, *fwd_ptr, *fwd_l, *fwd_k, *bwd_row, *bwd_ptr, *bwd_u, *bwd_k;
    // per-member state
    double *u, *uprev, *z, *du, *ur, *br, *xs;
    double *storage, *level, *factor, *area, *dlevel, *dfactor, *darea;
    double *precipitation, *surface_runoff, *potential_evaporation, *drainage, *infiltration;
    double *cum_precipitation, *cum_surface_runoff, *cum_drainage, *exact;
    double *fb_rate, *fb_cum;
    double *frp, *fro;  // current_flow_rate_pump / outlet (control caches)
    double *jv, *wv, *lu;
    double *norm_part, *ndz, *ndz_prev, *eta;
    int *nstat, *status, *active_count;
    double* storage_prev;  // limiter memory (basin.storage_prev, frozen: SURVEY row 22)
    double* level_prev;
    // DiscreteControl (shared structure, per-member truth and control states)
    const int *dc_cond_ptr, *dc_csv_ptr, *dc_sv_kind, *dc_sv_idx, *dc_logic_ptr, *dc_logic;
    const double *dc_sv_weight, *dc_thr_hi, *dc_thr_lo, *dc_sv_const;
    const int *pump_dc, *pump_cs_ptr, *outlet_dc, *outlet_cs_ptr, *lr_dc, *lr_cs_ptr, *trc_dc,
        *trc_cs_ptr, *trc_cs_table;
    const double *pump_cs_flow_rate, *pump_cs_min_flow_rate, *pump_cs_max_flow_rate,
        *pump_cs_min_upstream_level, *pump_cs_max_downstream_level;
    const double *outlet_cs_flow_rate, *outlet_cs_min_flow_rate, *outlet_cs_max_flow_rate,
        *outlet_cs_min_upstream_level, *outlet_cs_max_downstream_level;
    const double *lr_cs_resistance, *lr_cs_max_flow_rate;
    const int *mn_dc, *mn_cs_ptr, *pid_dc, *pid_cs_ptr;
    const double *mn_cs_length, *mn_cs_manning_n, *mn_cs_profile_width, *mn_cs_profile_slope;
    const double *pid_cs_target, *pid_cs_proportional, *pid_cs_integral, *pid_cs_derivative;
    int *dc_truth, *dc_state, *cnt_dc;
    // per-member sub-stepping state machine (rbs_step.cuh)
    int *mphase, *mjac, *miter, *mconv, *mneg, *mneg2, *mclamp, *mlast, *maction;
    int *cnt_iters, *cnt_sub, *cnt_rej;
    int* mstep;  // outer steps completed in the current window
    double *mh, *mhnom, *melapsed, *mhlast, *zlast;
    // forced window (rbs_advance_window): forcing slices [step][entity][member] per outer step of
    // the window (nullptr = the current array applies to every step) and per-step storage snapshots
    int fw_on;
    // outer step k of the window reads forcing slice (k + fw_phase) / fw_every (rbs_set_forcing_slicing:
    // e.g. daily forcing under hourly steps = one slice per 24 steps instead of 24 copies)
    int fw_every, fw_phase;
    long long fw_sb, fw_sf;  // elements per step: n_basin*B, n_fb*B
    const double *w_precipitation, *w_surface_runoff, *w_potential_evaporation, *w_drainage,
        *w_infiltration, *w_fb_rate;
    double* w_storage;
    double *w_flow, *ustep;  // mean flows per outer step [step][state][member]; u at its start
    double fw_dt;            // length of an outer step of the window
    // DiscreteControl record per member (callback.jl:698-716 `discrete_control.record`): entry r <
    // cnt_dc[m] = (time, node | state << 16, truth bits), the first dc_rec_cap changes are kept
    double win_t0, win_dt;   // start time and outer step of the running window
    int dc_rec_cap;
    double* dc_rec_t;
    int *dc_rec_ns, *dc_rec_bits;
    int* mroll;  // set by k_step_decide: the decision completed an outer step
    // 1 = the PidControl items run inside the phase-0 launch of k_links (allowed when no controller
    // reads a flow: all derivative gains zero, no ContinuousControl; decided on the host)
    int pid_in_links;
    // member-per-lane solve (rbs_lane.cuh): k_assemble_linear writes reduced right-hand-side row r
    // straight into row n_lu + lane_iperm[r] of `lu` (the solve's y column) instead of `br`
    const int* lane_iperm;
};

// Forcing array that applies to outer step `idx` of the window.
__device__ __forceinline__ const double* rbs_fw(const Net& n, const double* cur, const double* win,
                                                long long stride, int idx) {
    if (!(n.fw_on && win)) return cur;
    return win + stride * (n.fw_every > 1 ? (idx + n.fw_phase) / n.fw_every : idx);
}

// Member predicate of the step path: a kernel launched with `mphase != nullptr` only works on
// members whose phase equals `want`; with `mjac != nullptr` only on members whose Jacobian
// flag equals the kernel's JAC template argument.  The seam entry points pass nullptr.
struct Pred {
    const int* mphase;
    int want;
    const int* mjac;
};
template <bool JAC> __device__ __forceinline__ bool rbs_skip(const Pred& p, int m) {
    if (p.mphase && p.mphase[m] != p.want) return true;
    if (p.mjac && (p.mjac[m] != 0) != JAC) return true;
    return false;
}

#define RBS_TID (blockIdx.x * (long long)blockDim.x + threadIdx.x)
// resident blocks per SM the register allocation of the two big RHS kernels is bounded for
#ifndef RBS_BASIN_BLOCKS
#define RBS_BASIN_BLOCKS 6
#endif
#ifndef RBS_LINKS_BLOCKS
#define RBS_LINKS_BLOCKS 4
#endif
// Programmatic dependent launch (step-path kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): let the next kernel's blocks be scheduled
// as soon as every block of this one has started, and wait here until the previous kernel has
// completed and its writes are visible.  Takes the launch latency between the dozen short dependent
// kernels of a pass off the critical path; a no-op for kernels launched the ordinary way.
#define RBS_PDL() do { asm volatile("griddepcontrol.launch_dependents;"); \
                       asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
// member index of position j in the launch domain (compacted list of active members)
#define RBS_MEMBER(n, j) ((n).alist ? (n).alist[(j)] : (j))

struct RbsEnd { double h, f, dh, df; };

// Row of the control-state parameter table that applies to node k for member m, or -1 when the
// node is not under DiscreteControl (or no control state has been set yet): the shared static /
// time-dependent parameter applies.  Replaces the reference's in-place parameter swap
// (`set_control_params!`, core/src/callback.jl:765-786), which cannot be per ensemble member.
__device__ __forceinline__ int rbs_dc_row(const Net& n, const int* __restrict__ dc_of,
                                          const int* __restrict__ cs_ptr, int k, int m) {
    if (n.n_dc == 0) return -1;
    int d = dc_of[k];
    if (d < 0) return -1;
    int st = n.dc_state[(long long)d * n.B + m];
    return st < 0 ? -1 : cs_ptr[k] + st;
}
#define RBS_P(row, cs_arr, shared_arr, k) ((row) >= 0 ? (cs_arr)[(row)] : (shared_arr)[(k)])
// the same with a per-member override of the shared value (parameter ensembles)
#define RBS_PM(n, row, cs_arr, shared_arr, pm_arr, k, m) \
    ((row) >= 0 ? (cs_arr)[(row)] : ((pm_arr) ? (pm_arr)[(long long)(k) * (n).B + (m)] : (shared_arr)[(k)]))

template <bool JAC>
__device__ __forceinline__ RbsEnd rbs_fetch_end(const Net& n, int kind, int idx, int m) {
    RbsEnd e;
    if (kind == KIND_BASIN) {
        long long o = (long long)idx * n.B + m;
        e.h = n.level[o];
        e.f = n.factor[o];
        e.dh = JAC ? n.dlevel[o] : 0.0;
        e.df = JAC ? n.dfactor[o] : 0.0;
    } else if (kind == KIND_LEVEL_BOUNDARY) {
        e.h = n.pm_lb_level ? n.pm_lb_level[(long long)idx * n.B + m] : n.lb_level[idx];
        e.f = 1.0; e.dh = 0.0; e.df = 0.0;
    } else {  // Terminal: bottomless pit (core/src/util.jl:184-187)
        e.h = -RBS_INF; e.f = 1.0; e.dh = 0.0; e.df = 0.0;
    }
    return e;
}

// ---- exact cumulative forcing up to the RHS time (solve.jl:137-150, 86-99) -------------------
__global__ void k_exact(Net n, double t, double tprev) {
    RBS_PDL();
    long long tid = RBS_TID;
    int b = (int)(tid / n.NA);
    if (b >= n.n_basin) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)b * n.NA));
    long long o = (long long)b * n.B + m;
    double dt = t - tprev;
    double ccp = n.cum_precipitation[o] + n.fixed_area[b] * n.precipitation[o] * dt;
    double ccr = n.cum_surface_runoff[o] + dt * n.surface_runoff[o];
    double ccd = n.cum_drainage[o] + dt * n.drainage[o];
    double ex = (ccp + ccr) + ccd;
    for (int k = n.bfb_ptr[b]; k < n.bfb_ptr[b + 1]; k++) {
        long long f = (long long)n.bfb_idx[k] * n.B + m;
        ex += n.fb_cum[f] + n.fb_rate[f] * dt;
    }
    n.exact[o] = ex;
}

// ---- u_reduced = A * x (reduce_state!, util.jl:745-791) ---------------------------------------
// MODE 0: x = a;  MODE 1: x = a + b (uprev + z);  MODE 2: x = (dt*a - b) * inv_dt (Newton rhs)
template <int MODE>
__device__ __forceinline__ double rbs_red_in(const double* a, const double* b, long long o,
                                             double dt, double inv_dt) {
    if (MODE == 0) return a[o];
    if (MODE == 1) return a[o] + b[o];
    return (dt * a[o] - b[o]) * inv_dt;
}

template <int MODE>
__global__ void k_reduce(Net n, const double* __restrict__ a, const double* __restrict__ b,
                         double* __restrict__ out, double dt, double inv_dt) {
    RBS_PDL();
    long long tid = RBS_TID;
    int r = (int)(tid / n.NA);
    if (r >= n.n_reduced) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)r * n.NA));
    if (r < n.n_basin) {
        double acc = 0.0;
        for (int k = n.inc_ptr[r]; k < n.inc_ptr[r + 1]; k++) {
            double v = rbs_red_in<MODE>(a, b, (long long)n.inc_state[k] * n.B + m, dt, inv_dt);
            acc = n.inc_sign[k] > 0.0 ? acc + v : acc - v;
        }
        acc -= rbs_red_in<MODE>(a, b, (long long)(n.off_evap + r) * n.B + m, dt, inv_dt);
        acc -= rbs_red_in<MODE>(a, b, (long long)(n.off_infil + r) * n.B + m, dt, inv_dt);
        out[(long long)r * n.B + m] = acc;
    } else {
        int p = r - n.n_basin;
        out[(long long)r * n.B + m] =
            rbs_red_in<MODE>(a, b, (long long)(n.off_integral + p) * n.B + m, dt, inv_dt);
    }
}

// ---- basin properties + vertical fluxes (solve.jl:120-227) -------------------------------------
// SRC 0: u_reduced given in `ur`;  SRC 1: reduce_state! of (uprev + z) fused in;  SRC 2: of u.
// Grid: n_reduced x B (rows >= n_basin only forward the PID integral state into n.ur).
template <int SRC> __device__ __forceinline__ double rbs_state_in(const Net& n, long long o) {
    return SRC == 1 ? n.uprev[o] + n.z[o] : n.u[o];
}
template <bool JAC, int SRC>
__device__ __forceinline__ void rbs_basin_item(const Net& n, const double* ur, bool write_du,
                                               int flag_negative, int b, int m) {
    // flag_negative: 0 none; 1 = proposed state (sets mneg); 2 = clamped state (only members the
    // limiter changed, sets mneg2)
    if (flag_negative == 2 && n.mclamp[m] == 0) return;
    long long o = (long long)b * n.B + m;
    if (b >= n.n_basin) {
        if (SRC != 0) n.ur[o] = rbs_state_in<SRC>(n, (long long)(n.off_integral + b - n.n_basin) * n.B + m);
        return;
    }
    double urv;
    if (SRC == 0) {
        urv = ur[o];
    } else {
        double acc = 0.0;
        for (int k = n.inc_ptr[b]; k < n.inc_ptr[b + 1]; k++) {
            double v = rbs_state_in<SRC>(n, (long long)n.inc_state[k] * n.B + m);
            acc = n.inc_sign[k] > 0.0 ? acc + v : acc - v;
        }
        acc -= rbs_state_in<SRC>(n, (long long)(n.off_evap + b) * n.B + m);
        acc -= rbs_state_in<SRC>(n, (long long)(n.off_infil + b) * n.B + m);
        urv = acc;
        n.ur[o] = acc;
    }
    double S = (n.storage0[b] + urv) + n.exact[o];
    double dphi;
    double phi = rbs_reduction_factor(S, n.low_storage_threshold[b], dphi);
    int p0 = n.prof_ptr[b], np = n.prof_ptr[b + 1] - p0;
    RbsBasinEval e = rbs_basin_eval(S, np, n.prof_level + p0, n.prof_dSdh + p0,
                                    n.prof_dSdh_slope + p0, n.prof_cumS + p0, n.prof_area + p0,
                                    n.prof_area_slope + p0);
    // Newton iterates of the step path (SRC 1 with the Jacobian): the links only read level, factor
    // and their partials; storage / area are written for the basins a controller observes.  The
    // finalize half and the seam calls write everything.
    const bool lean = JAC && SRC == 1 && !n.basin_obs[b];
    n.factor[o] = phi;
    n.level[o] = e.level;
    if (!lean) {
        n.storage[o] = S;
        n.area[o] = e.area;
    }
    if (flag_negative && S < 0.0) atomicOr(flag_negative == 2 ? &n.mneg2[m] : &n.mneg[m], 1);  // integer flag only
    if (JAC) {
        n.dlevel[o] = e.dlevel;
        n.dfactor[o] = dphi;
        if (!lean) n.darea[o] = e.darea;
    }
    if (write_du) {
        const int ws = n.fw_on ? n.mstep[m] : 0;
        double E = rbs_fw(n, n.potential_evaporation, n.w_potential_evaporation, n.fw_sb, ws)[o];
        double I = rbs_fw(n, n.infiltration, n.w_infiltration, n.fw_sb, ws)[o];
        n.du[(long long)(n.off_evap + b) * n.B + m] = e.area * phi * E;
        n.du[(long long)(n.off_infil + b) * n.B + m] = phi * I;
        if (JAC) {
            n.jv[(long long)n.jrow_ptr[n.off_evap + b] * n.B + m] = (e.darea * phi + e.area * dphi) * E;
            n.jv[(long long)n.jrow_ptr[n.off_infil + b] * n.B + m] = dphi * I;
        }
    }
}

template <bool JAC, int SRC>
__global__ void __launch_bounds__(256, RBS_BASIN_BLOCKS) k_basin(Net n, const double* __restrict__ ur, bool write_du, Pred pred,
                        int flag_negative) {
    RBS_PDL();
    long long tid = RBS_TID;
    int b = (int)(tid / n.NA);
    if (b >= n.n_reduced) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)b * n.NA));
    if (rbs_skip<JAC>(pred, m)) return;
    rbs_basin_item<JAC, SRC>(n, ur, write_du, flag_negative, b, m);
}

// ---- one user-demand inflow link (solve.jl:385-397) ---------------------------------------------
template <bool JAC>
__device__ __forceinline__ double rbs_ud_link(const Net& n, int k, int m, double& dq) {
    int i = n.ud_of_link[k];
    int s = n.off_ud_in + k;
    RbsEnd a = rbs_fetch_end<JAC>(n, n.src_kind[s], n.src_idx[s], m);
    int n_links = n.ud_link_ptr[i + 1] - n.ud_link_ptr[i];
    double equal_split = n_links == 0 ? 0.0 : n.ud_total_demand[i] / n_links;
    double alloc = n.ud_link_alloc[k];
    double qt = isinf(alloc) ? equal_split : alloc;
    double rp;
    double r = rbs_reduction_factor(a.h - n.ud_min_level[i], n.ldt, rp);
    if (JAC) dq = qt * a.df * r + qt * a.f * (rp * a.dh);
    return qt * a.f * r;
}

// ---- connector flows (formulate_flows!, solve.jl:811-834) ---------------------------------------
// PHASE 0: every connector state whose flow does not depend on a controller;
// PHASE 1/2: the pumps/outlets listed in `list` (controlled by ContinuousControl / PidControl).
// flow (and its partials) of connector state s for member m
template <bool JAC>
__device__ __forceinline__ void rbs_link_item(const Net& n, int s, int m, int phase) {
    int ct = n.conn_type[s], k = n.conn_idx[s];
    long long so = (long long)s * n.B + m;
    double q = 0.0, dq_a = 0.0, dq_b = 0.0;
    if (ct == CT_UD_OUT) {
        // outflow = (sum of the node's link flows) * return_factor (solve.jl:384-401)
        double total = 0.0;
        for (int l = n.ud_link_ptr[k]; l < n.ud_link_ptr[k + 1]; l++) {
            double dq;
            double qk = rbs_ud_link<JAC>(n, l, m, dq);
            total += qk;
            if (JAC && n.ud_out_jpos[l] >= 0)
                n.jv[(long long)n.ud_out_jpos[l] * n.B + m] = dq * n.ud_return_factor[k];
        }
        n.du[so] = total * n.ud_return_factor[k];
        return;
    }
    if (ct == CT_UD_IN) {
        q = rbs_ud_link<JAC>(n, k, m, dq_a);
        n.du[so] = q;
        if (JAC && n.jpos_src[s] >= 0) n.jv[(long long)n.jpos_src[s] * n.B + m] = dq_a;
        return;
    }
    bool controlled = false;
    double g_fr = 0.0;
    if (ct == CT_PUMP || ct == CT_OUTLET) {
        bool outlet = ct == CT_OUTLET;
        int control = outlet ? n.outlet_control_type[k] : n.pump_control_type[k];
        if (control != phase) {
            // controllers read du of flows that are not formulated yet: those must read 0
            // (water_balance! zeroes du first, core/src/solve.jl:54); not when the PID item of the
            // same launch writes the flow and nobody reads it in between
            if (phase == 0 && !(control == CONTROL_PID && n.pid_in_links)) n.du[so] = 0.0;
            return;
        }
        controlled = control != CONTROL_NONE;
        RbsEnd a = rbs_fetch_end<JAC>(n, n.src_kind[s], n.src_idx[s], m);
        RbsEnd b = rbs_fetch_end<JAC>(n, n.dst_kind[s], n.dst_idx[s], m);
        int row = outlet ? rbs_dc_row(n, n.outlet_dc, n.outlet_cs_ptr, k, m)
                         : rbs_dc_row(n, n.pump_dc, n.pump_cs_ptr, k, m);
        double fr = controlled ? (outlet ? n.fro : n.frp)[(long long)k * n.B + m]
                               : (outlet ? RBS_PM(n, row, n.outlet_cs_flow_rate, n.outlet_flow_rate, n.pm_outlet_flow_rate, k, m)
                                         : RBS_PM(n, row, n.pump_cs_flow_rate, n.pump_flow_rate, n.pm_pump_flow_rate, k, m));
        q = rbs_pump_outlet_flow(
            fr, a.h, b.h, a.f, a.dh, b.dh, a.df,
            outlet ? RBS_P(row, n.outlet_cs_min_flow_rate, n.outlet_min_flow_rate, k)
                   : RBS_P(row, n.pump_cs_min_flow_rate, n.pump_min_flow_rate, k),
            outlet ? RBS_P(row, n.outlet_cs_max_flow_rate, n.outlet_max_flow_rate, k)
                   : RBS_P(row, n.pump_cs_max_flow_rate, n.pump_max_flow_rate, k),
            (outlet ? n.outlet_flow_demand : n.pump_flow_demand)[k],
            outlet ? RBS_P(row, n.outlet_cs_min_upstream_level, n.outlet_min_upstream_level, k)
                   : RBS_P(row, n.pump_cs_min_upstream_level, n.pump_min_upstream_level, k),
            outlet ? RBS_P(row, n.outlet_cs_max_downstream_level, n.outlet_max_downstream_level, k)
                   : RBS_P(row, n.pump_cs_max_downstream_level, n.pump_max_downstream_level, k),
            n.ldt, outlet, g_fr, dq_a, dq_b);
    } else {
        RbsEnd a = rbs_fetch_end<JAC>(n, n.src_kind[s], n.src_idx[s], m);
        RbsEnd b = rbs_fetch_end<JAC>(n, n.dst_kind[s], n.dst_idx[s], m);
        if (ct == CT_LR) {
            int row = rbs_dc_row(n, n.lr_dc, n.lr_cs_ptr, k, m);
            q = rbs_lr_flow(a.h, b.h, a.f, b.f, a.dh, b.dh, a.df, b.df,
                            RBS_PM(n, row, n.lr_cs_resistance, n.lr_resistance, n.pm_lr_resistance, k, m),
                            RBS_P(row, n.lr_cs_max_flow_rate, n.lr_max_flow_rate, k), dq_a, dq_b);
        } else if (ct == CT_MANNING) {
            int row = rbs_dc_row(n, n.mn_dc, n.mn_cs_ptr, k, m);
            q = rbs_manning_flow(a.h, b.h, a.f, b.f, a.dh, b.dh, a.df, b.df,
                                 RBS_P(row, n.mn_cs_length, n.mn_length, k),
                                 RBS_PM(n, row, n.mn_cs_manning_n, n.mn_manning_n, n.pm_mn_manning_n, k, m),
                                 RBS_P(row, n.mn_cs_profile_width, n.mn_profile_width, k),
                                 RBS_P(row, n.mn_cs_profile_slope, n.mn_profile_slope, k),
                                 n.mn_upstream_bottom[k], n.mn_downstream_bottom[k], dq_a, dq_b, JAC);
        } else {  // CT_TRC
            int row = rbs_dc_row(n, n.trc_dc, n.trc_cs_ptr, k, m);
            int tb = row >= 0 ? n.trc_cs_table[row] : n.trc_table_idx[k];
            int t0 = n.tab_ptr[tb], tn = n.tab_ptr[tb + 1] - t0;
            double dQ;
            double Q = rbs_pchip_eval(a.h, tn, n.tab_t + t0, n.tab_u + t0, n.tab_du + t0,
                                      n.tab_c1 + t0, n.tab_c2 + t0, n.tab_slope_left[tb],
                                      n.tab_slope_right[tb], dQ);
            q = rbs_trc_flow(a.h, b.h, a.f, a.dh, b.dh, a.df, Q, dQ,
                             n.trc_max_downstream_level[k], n.ldt, dq_a, dq_b);
        }
    }
    n.du[so] = q;
    if (JAC) {
        if (controlled) {
            // the controller left d(flow_rate)/dS in this row's slots: chain through dq/dflow_rate
            for (int e = n.jrow_ptr[s]; e < n.jrow_ptr[s + 1]; e++) n.jv[(long long)e * n.B + m] *= g_fr;
            if (n.jpos_src[s] >= 0) n.jv[(long long)n.jpos_src[s] * n.B + m] += dq_a;
            if (n.jpos_dst[s] >= 0) n.jv[(long long)n.jpos_dst[s] * n.B + m] += dq_b;
        } else {
            if (n.jpos_src[s] >= 0) n.jv[(long long)n.jpos_src[s] * n.B + m] = dq_a;
            if (n.jpos_dst[s] >= 0) n.jv[(long long)n.jpos_dst[s] * n.B + m] = dq_b;
        }
    }
}

template <bool JAC>
__device__ __forceinline__ void rbs_pid_item(const Net& n, const double* ur, int fuse_link, int i, int m);

template <bool JAC>
__global__ void __launch_bounds__(256, RBS_LINKS_BLOCKS) k_links(Net n, int phase, const int* __restrict__ list, int n_list, Pred pred) {
    RBS_PDL();
    long long tid = RBS_TID;
    int li = (int)(tid / n.NA);
    int count = phase == 0 ? n.n_conn + (n.pid_in_links ? n.n_pid : 0) : n_list;
    if (li >= count) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)li * n.NA));
    if (rbs_skip<JAC>(pred, m)) return;
    if (phase == 0 && li >= n.n_conn) {
        // PidControl item with its controlled pump / outlet (what k_pid does, one launch less)
        rbs_pid_item<JAC>(n, n.ur, 1, li - n.n_conn, m);
        return;
    }
    rbs_link_item<JAC>(n, phase == 0 ? li : list[li], m, phase);
}

// ---- ContinuousControl (solve.jl:101-113, callback.jl:757-763) ----------------------------------
template <bool JAC>
__device__ __forceinline__ void rbs_cc_item(const Net& n, int i, int m) {
    double value = 0.0;
    for (int j = n.cc_sv_ptr[i]; j < n.cc_sv_ptr[i + 1]; j++) {
        int kind = n.cc_sv_kind[j], idx = n.cc_sv_idx[j];
        double x;
        if (kind == 0) x = n.level[(long long)idx * n.B + m];
        else if (kind == 1) x = n.storage[(long long)idx * n.B + m];
        else if (kind == 2) x = n.du[(long long)idx * n.B + m];
        else x = n.cc_sv_const[j];
        value += n.cc_sv_weight[j] * x;
    }
    int tb = n.cc_tab_offset + i;
    int t0 = n.tab_ptr[tb], tn = n.tab_ptr[tb + 1] - t0;
    double df;
    double out = rbs_pchip_eval(value, tn, n.tab_t + t0, n.tab_u + t0, n.tab_du + t0, n.tab_c1 + t0,
                                n.tab_c2 + t0, n.tab_slope_left[tb], n.tab_slope_right[tb], df);
    (n.cc_target_is_outlet[i] ? n.fro : n.frp)[(long long)n.cc_target_idx[i] * n.B + m] = out;
    if (JAC) {
        int row = n.cc_target_state[i];
        for (int e = n.jrow_ptr[row]; e < n.jrow_ptr[row + 1]; e++) n.jv[(long long)e * n.B + m] = 0.0;
        for (int q = n.cc_term_ptr[i]; q < n.cc_term_ptr[i + 1]; q++) {
            double w = n.cc_sv_weight[n.cc_term_sv[q]];
            int kind = n.cc_term_kind[q], src = n.cc_term_src[q];
            double dx = kind == 0 ? n.dlevel[(long long)src * n.B + m]
                                  : (kind == 1 ? 1.0 : n.jv[(long long)src * n.B + m]);
            n.jv[(long long)n.cc_term_dst[q] * n.B + m] += df * (w * dx);
        }
    }
}

template <bool JAC>
__global__ void k_continuous_control(Net n, Pred pred) {
    RBS_PDL();
    long long tid = RBS_TID;
    int i = (int)(tid / n.NA);
    if (i >= n.n_cc) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)i * n.NA));
    if (rbs_skip<JAC>(pred, m)) return;
    rbs_cc_item<JAC>(n, i, m);
}

// ---- PidControl (solve.jl:229-346) ---------------------------------------------------------------
// fuse_link: also formulate the flow of the controlled pump / outlet in the same thread (allowed
// when no controller reads the flow of a PID-controlled node, so that the launch order between
// the controllers and the controlled links does not matter: checked on the host)
template <bool JAC>
__device__ __forceinline__ void rbs_pid_item(const Net& n, const double* ur, int fuse_link, int i, int m) {
    int L = n.pid_listen_basin[i];
    long long lo = (long long)L * n.B + m;
    // a DiscreteControl state replaces the (time-dependent) parameters by its constants
    // (PidControl control_mapping, read.jl:328-427)
    const int crow = rbs_dc_row(n, n.pid_dc, n.pid_cs_ptr, i, m);
    double err = RBS_P(crow, n.pid_cs_target, n.pid_target, i) - n.level[lo];
    n.du[(long long)(n.off_integral + i) * n.B + m] = err;
    double K_p = RBS_P(crow, n.pid_cs_proportional, n.pid_proportional, i),
           K_i = RBS_P(crow, n.pid_cs_integral, n.pid_integral, i),
           K_d = RBS_P(crow, n.pid_cs_derivative, n.pid_derivative, i);
    double integral = ur[(long long)(n.n_basin + i) * n.B + m];
    double area = 1.0, D = 1.0;
    if (K_d != 0.0) {
        area = n.area[lo];
        D = 1.0 - K_d / area;
    }
    double flow_rate = 0.0;
    if (K_p != 0.0) flow_rate += K_p * err / D;
    if (K_i != 0.0) flow_rate += K_i * integral / D;
    double dS = 0.0, N = 0.0;
    if (K_d != 0.0) {
        // formulate_dstorage_wrt_time (solve.jl:321-346) on the flows formulated so far
        for (int k = n.pid_term_ptr[i]; k < n.pid_term_ptr[i + 1]; k++) {
            double v = n.du[(long long)n.pid_term_state[k] * n.B + m];
            dS = n.pid_term_sign[k] > 0.0 ? dS + v : dS - v;
        }
        // FlowBoundary inflows enter with their instantaneous rate (graph.jl:346-366)
        // inflow order: neighbour list order is (connector states..., boundaries) — the
        // boundary rates are added after the connector flows (rounding-level difference only)
        const int ws = n.fw_on ? n.mstep[m] : 0;
        const double* fbr = rbs_fw(n, n.fb_rate, n.w_fb_rate, n.fw_sf, ws);
        for (int k = n.bfb_ptr[L]; k < n.bfb_ptr[L + 1]; k++)
            dS += fbr[(long long)n.bfb_idx[k] * n.B + m];
        dS += n.fixed_area[L] * rbs_fw(n, n.precipitation, n.w_precipitation, n.fw_sb, ws)[lo];
        dS += rbs_fw(n, n.surface_runoff, n.w_surface_runoff, n.fw_sb, ws)[lo];
        dS += rbs_fw(n, n.drainage, n.w_drainage, n.fw_sb, ws)[lo];
        dS -= n.du[(long long)(n.off_evap + L) * n.B + m];
        dS -= n.du[(long long)(n.off_infil + L) * n.B + m];
        N = 0.0 - dS / area;  // dtarget = 0 for a ConstantInterpolation target (solve.jl:298-300)
        flow_rate += K_d * N / D;
    }
    (n.pid_target_is_outlet[i] ? n.fro : n.frp)[(long long)n.pid_target_idx[i] * n.B + m] = flow_rate;
    if (JAC) {
        double dh = n.dlevel[lo];
        // integral row: d(err)/dS_L
        n.jv[(long long)n.pid_int_jpos[i] * n.B + m] = -dh;
        int row = n.pid_target_state[i];
        for (int e = n.jrow_ptr[row]; e < n.jrow_ptr[row + 1]; e++) n.jv[(long long)e * n.B + m] = 0.0;
        double dA = K_d != 0.0 ? n.darea[lo] : 0.0;
        double dD = K_d != 0.0 ? K_d * dA / (area * area) : 0.0;  // dD/dS_L
        double d_listen = 0.0;
        if (K_p != 0.0) d_listen += K_p * ((-dh) * D - err * dD) / (D * D);
        if (K_i != 0.0) {
            d_listen += -K_i * integral * dD / (D * D);
            n.jv[(long long)n.pid_jpos_integral[i] * n.B + m] = K_i / D;
        }
        if (K_d != 0.0) {
            // d(N)/dS_c = -(d(dS)/dS_c * area - dS * dA[c == L]) / area^2
            double coef = K_d / D;
            for (int k = n.pid_jterm_ptr[i]; k < n.pid_jterm_ptr[i + 1]; k++) {
                double v = n.pid_jterm_sign[k] * n.jv[(long long)n.pid_jterm_src[k] * n.B + m];
                n.jv[(long long)n.pid_jterm_dst[k] * n.B + m] += coef * (-(v) / area);
            }
            double dvert = n.jv[(long long)n.jrow_ptr[n.off_evap + L] * n.B + m] +
                           n.jv[(long long)n.jrow_ptr[n.off_infil + L] * n.B + m];
            d_listen += coef * (dvert / area);       // -(-dvert)/area
            d_listen += coef * (dS * dA / (area * area));
            d_listen += -K_d * N * dD / (D * D);
        }
        n.jv[(long long)n.pid_jpos_listen[i] * n.B + m] += d_listen;
    }
    if (fuse_link) rbs_link_item<JAC>(n, n.pid_target_state[i], m, CONTROL_PID);
}
template <bool JAC>
__global__ void k_pid(Net n, const double* __restrict__ ur, Pred pred, int fuse_link) {
    RBS_PDL();
    long long tid = RBS_TID;
    int i = (int)(tid / n.NA);
    if (i >= n.n_pid) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)i * n.NA));
    if (rbs_skip<JAC>(pred, m)) return;
    rbs_pid_item<JAC>(n, ur, fuse_link, i, m);
}

// ---- W_inner = A*J_i - I/gamma (calc_J_inner! + jacobian2W!, differentiation.jl:153-271) --------
// One W entry: the ordered signed sum of the J_intermediate entries that land on it, minus
// 1/gamma on the diagonal.
__device__ __forceinline__ double rbs_w_entry(const Net& n, int e, int m, double inv_gamma) {
    double acc = 0.0;
    for (int k = n.wsrc_ptr[e]; k < n.wsrc_ptr[e + 1]; k++) {
        double v = n.jv[(long long)n.wsrc_pos[k] * n.B + m];
        acc = n.wsrc_sign[k] > 0.0 ? acc + v : acc - v;
    }
    if (n.wdiag_flag[e]) acc = acc - inv_gamma;
    return acc;
}
__global__ void k_assemble_w(Net n, double inv_gamma) {
    RBS_PDL();
    long long tid = RBS_TID;
    int e = (int)(tid / n.NA);
    if (e >= n.nnz_w) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)e * n.NA));
    n.wv[(long long)e * n.B + m] = rbs_w_entry(n, e, m, inv_gamma);
}

// gamma of member m: shared (seam calls) or 1/h of the member's current sub-step
__device__ __forceinline__ double rbs_inv_gamma(const Net& n, int m, double inv_gamma) {
    return inv_gamma > 0.0 ? inv_gamma : 1.0 / n.mh[m];
}

// ---- static-pattern LU: one entry of L or U (see ribasim_b200/symbolic.py) ----------------------
// FROM_J: W is assembled on the fly from J_intermediate (step path, no W array traffic).
template <bool FROM_J>
__device__ __forceinline__ void rbs_lu_entry(const Net& n, int e, int m, double inv_gamma) {
    int src = n.lu_wsrc[e];
    double v = 0.0;
    if (src >= 0) v = FROM_J ? rbs_w_entry(n, src, m, inv_gamma) : n.wv[(long long)src * n.B + m];
    for (int q = n.lu_prod_ptr[e]; q < n.lu_prod_ptr[e + 1]; q++)
        v -= n.lu[(long long)n.lu_prod_l[q] * n.B + m] * n.lu[(long long)n.lu_prod_u[q] * n.B + m];
    int pv = n.lu_pivot[e];
    if (pv >= 0) v /= n.lu[(long long)pv * n.B + m];
    n.lu[(long long)e * n.B + m] = v;
}

// wide mode: one launch per dependency level, grid over (entries of the level) x members
template <bool FROM_J>
__global__ void k_lu_level(Net n, int e0, int e1, double inv_gamma, Pred pred) {
    RBS_PDL();
    long long tid = RBS_TID;
    int e = e0 + (int)(tid / n.NA);
    if (e >= e1) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)(e - e0) * n.NA));
    if (rbs_skip<true>(pred, m)) return;
    rbs_lu_entry<FROM_J>(n, e, m, FROM_J ? rbs_inv_gamma(n, m, inv_gamma) : 0.0);
}

// tile mode: a block owns a tile of members for the whole factorisation; levels are separated
// by __syncthreads() only (no grid-wide dependency because members are independent).
template <bool FROM_J>
__global__ void k_lu_tile(Net n, const int* __restrict__ level_ptr, int n_levels, int tile,
                          double inv_gamma, Pred pred) {
    RBS_PDL();
    int m0 = blockIdx.x * tile;
    int tm = min(tile, n.NA - m0);
    if (tm <= 0) return;
    for (int lv = 0; lv < n_levels; lv++) {
        int e0 = level_ptr[lv], e1 = level_ptr[lv + 1];
        int work = (e1 - e0) * tm;
        for (int w = threadIdx.x; w < work; w += blockDim.x) {
            int e = e0 + w / tm;
            int m = RBS_MEMBER(n, m0 + w % tm);
            if (rbs_skip<true>(pred, m)) continue;
            rbs_lu_entry<FROM_J>(n, e, m, FROM_J ? rbs_inv_gamma(n, m, inv_gamma) : 0.0);
        }
        __syncthreads();
    }
}

// ---- triangular solves ----------------------------------------------------------------------------
// Right-hand side of reduced row r: either b[r] (seam) or, on the step path, row r of
// A * ((h*du - z)/h) computed on the fly (reduce_state!(b), differentiation.jl:321-325).
template <bool FROM_DU>
__device__ __forceinline__ double rbs_solve_rhs(const Net& n, const double* __restrict__ b, int r, int m) {
    if (!FROM_DU) return b[(long long)r * n.B + m];
    double h = n.mh[m], inv_h = 1.0 / h;
    if (r < n.n_basin) {
        double acc = 0.0;
        for (int k = n.inc_ptr[r]; k < n.inc_ptr[r + 1]; k++) {
            double v = rbs_red_in<2>(n.du, n.z, (long long)n.inc_state[k] * n.B + m, h, inv_h);
            acc = n.inc_sign[k] > 0.0 ? acc + v : acc - v;
        }
        acc -= rbs_red_in<2>(n.du, n.z, (long long)(n.off_evap + r) * n.B + m, h, inv_h);
        acc -= rbs_red_in<2>(n.du, n.z, (long long)(n.off_infil + r) * n.B + m, h, inv_h);
        return acc;
    }
    return rbs_red_in<2>(n.du, n.z, (long long)(n.off_integral + r - n.n_basin) * n.B + m, h, inv_h);
}
template <bool FROM_DU>
__global__ void k_solve_load(Net n, const double* __restrict__ b, Pred pred) {  // xs = P b
    RBS_PDL();
    long long tid = RBS_TID;
    int i = (int)(tid / n.NA);
    if (i >= n.n_reduced) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)i * n.NA));
    if (rbs_skip<true>(pred, m)) return;
    n.xs[(long long)i * n.B + m] = rbs_solve_rhs<FROM_DU>(n, b, n.perm[i], m);
}
__device__ __forceinline__ void rbs_fwd_row(const Net& n, int pos, int m) {
    int i = n.fwd_row[pos];
    double v = n.xs[(long long)i * n.B + m];
    for (int q = n.fwd_ptr[pos]; q < n.fwd_ptr[pos + 1]; q++)
        v -= n.lu[(long long)n.fwd_l[q] * n.B + m] * n.xs[(long long)n.fwd_k[q] * n.B + m];
    n.xs[(long long)i * n.B + m] = v;
}
__device__ __forceinline__ void rbs_bwd_row(const Net& n, int pos, int m) {
    int i = n.bwd_row[pos];
    double v = n.xs[(long long)i * n.B + m];
    for (int q = n.bwd_ptr[pos]; q < n.bwd_ptr[pos + 1]; q++)
        v -= n.lu[(long long)n.bwd_u[q] * n.B + m] * n.xs[(long long)n.bwd_k[q] * n.B + m];
    n.xs[(long long)i * n.B + m] = v / n.lu[(long long)n.lu_diag[i] * n.B + m];
}
__global__ void k_solve_level(Net n, int p0, int p1, int backward, Pred pred) {
    RBS_PDL();
    long long tid = RBS_TID;
    int pos = p0 + (int)(tid / n.NA);
    if (pos >= p1) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)(pos - p0) * n.NA));
    if (rbs_skip<true>(pred, m)) return;
    if (backward) rbs_bwd_row(n, pos, m); else rbs_fwd_row(n, pos, m);
}
template <bool FROM_DU>
__global__ void k_solve_tile(Net n, const double* __restrict__ b, double* __restrict__ x,
                             const int* __restrict__ fwd_level_ptr, int n_fwd,
                             const int* __restrict__ bwd_level_ptr, int n_bwd, int tile, Pred pred) {
    RBS_PDL();
    int m0 = blockIdx.x * tile;
    int tm = min(tile, n.NA - m0);
    if (tm <= 0) return;
    for (int w = threadIdx.x; w < n.n_reduced * tm; w += blockDim.x) {
        int i = w / tm, m = RBS_MEMBER(n, m0 + w % tm);
        if (rbs_skip<true>(pred, m)) continue;
        n.xs[(long long)i * n.B + m] = rbs_solve_rhs<FROM_DU>(n, b, n.perm[i], m);
    }
    __syncthreads();
    for (int lv = 1; lv < n_fwd; lv++) {  // level 0 rows have no dependencies
        int p0 = fwd_level_ptr[lv], p1 = fwd_level_ptr[lv + 1];
        for (int w = threadIdx.x; w < (p1 - p0) * tm; w += blockDim.x) {
            int m = RBS_MEMBER(n, m0 + w % tm);
            if (!rbs_skip<true>(pred, m)) rbs_fwd_row(n, p0 + w / tm, m);
        }
        __syncthreads();
    }
    for (int lv = 0; lv < n_bwd; lv++) {
        int p0 = bwd_level_ptr[lv], p1 = bwd_level_ptr[lv + 1];
        for (int w = threadIdx.x; w < (p1 - p0) * tm; w += blockDim.x) {
            int m = RBS_MEMBER(n, m0 + w % tm);
            if (!rbs_skip<true>(pred, m)) rbs_bwd_row(n, p0 + w / tm, m);
        }
        __syncthreads();
    }
    for (int w = threadIdx.x; w < n.n_reduced * tm; w += blockDim.x) {
        int i = w / tm, m = RBS_MEMBER(n, m0 + w % tm);
        if (rbs_skip<true>(pred, m)) continue;
        x[(long long)n.perm[i] * n.B + m] = n.xs[(long long)i * n.B + m];
    }
}
__global__ void k_solve_store(Net n, double* __restrict__ x, Pred pred) {  // x = P^T xs
    RBS_PDL();
    long long tid = RBS_TID;
    int i = (int)(tid / n.NA);
    if (i >= n.n_reduced) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)i * n.NA));
    if (rbs_skip<true>(pred, m)) return;
    x[(long long)n.perm[i] * n.B + m] = n.xs[(long long)i * n.B + m];
}

// ---- step path: fused factorisation + substitution of one member tile -----------------------------
// A block owns TILE members of the launch domain (TILE = 32: a warp handles one L/U entry or one
// solution row for 32 consecutive members, every global access is one coalesced 256 B request).
//   * the whole elimination / substitution schedule (index arrays, packed to 16 bit when the
//     network allows) is staged in shared memory once per block, so the ~100 dependent levels
//     never wait on an index load;
//   * the L/U factors stay in the global `lu` array: each block touches n_lu*TILE*8 B of it,
//     written and re-read within microseconds, i.e. it lives in the 126 MB L2; all operands of
//     an entry (its own slot, the pivot, the product pairs) are independent loads issued together;
//   * the solution vectors live in shared memory ([n_reduced][TILE]);
//   * W = A J - I/h is assembled on the fly from J_intermediate (no W array), the right-hand side
//     A*((h du - z)/h) is gathered on the fly (no b array).
// Members that reuse their factors (frozen-Newton iterations >= 2) skip the elimination.
struct LinSched {          // offsets (in index elements) into the packed schedule blob
    int prod_ptr, pivot, prod_l, prod_u, diag, perm;
    int fwd_row, fwd_ptr, fwd_l, fwd_k, bwd_row, bwd_ptr, bwd_u, bwd_k;
    int lvl_lu, lvl_fwd, lvl_bwd;  // level pointers
    int total;             // number of index elements
    int n_levels, n_fwd, n_bwd;
};

template <int TILE, class IT>
__global__ void __launch_bounds__(1024)
k_newton_linear(Net n, LinSched sc, const IT* __restrict__ blob, const int* __restrict__ lsrc_ptr,
                const int* __restrict__ lsrc_code, double* __restrict__ x) {
    RBS_PDL();
    extern __shared__ double smem[];
    double* xs_s = smem;                               // [n_reduced][TILE]
    IT* ix = (IT*)(xs_s + (size_t)n.n_reduced * TILE);
    __shared__ int mem_s[TILE], fac_s[TILE];           // member index (-1 = idle), refresh flag
    __shared__ double invh_s[TILE];
    __shared__ int any_fac;
    const IT NONE = (IT)~(IT)0;
    constexpr int LOG = TILE == 32 ? 5 : (TILE == 16 ? 4 : (TILE == 8 ? 3 : (TILE == 4 ? 2 : (TILE == 2 ? 1 : 0))));
    int m0 = blockIdx.x * TILE;
    int tm = min(TILE, n.NA - m0);
    if (tm <= 0) return;
    if (threadIdx.x == 0) any_fac = 0;
    __syncthreads();
    if (threadIdx.x < TILE) {
        int j = threadIdx.x;
        int m = j < tm ? RBS_MEMBER(n, m0 + j) : -1;
        if (m >= 0 && n.mphase[m] != PHASE_NEWTON) m = -1;
        mem_s[j] = m;
        int f = m >= 0 && n.mjac[m] != 0;
        fac_s[j] = f;
        if (f) any_fac = 1;
        invh_s[j] = m >= 0 ? 1.0 / n.mh[m] : 0.0;
    }
    for (int w = threadIdx.x; w < sc.total; w += blockDim.x) ix[w] = blob[w];
    __syncthreads();
    const int j = threadIdx.x & (TILE - 1);
    const int lane_e = threadIdx.x >> LOG, stride_e = blockDim.x >> LOG;
    const int m = mem_s[j];
    const bool fac = fac_s[j] != 0;
    double* lu = n.lu + (m >= 0 ? m : 0);
    const long long B = n.B;
    if (any_fac) {
        // 1. initialise the factor slots from J (ordered signed sums; code < 0 = "- 1/h")
        if (fac) {
            double invh = invh_s[j];
            for (int e = lane_e; e < n.n_lu; e += stride_e) {
                double v = 0.0;
                for (int k = lsrc_ptr[e]; k < lsrc_ptr[e + 1]; k++) {
                    int code = lsrc_code[k];
                    if (code < 0) { v = v - invh; continue; }
                    double jvv = n.jv[(long long)(code >> 1) * B + m];
                    v = (code & 1) ? v - jvv : v + jvv;
                }
                lu[(long long)e * B] = v;
            }
        }
        __syncthreads();
        // 2. elimination, level by level
        const IT *pp = ix + sc.prod_ptr, *pv_ = ix + sc.pivot, *pl = ix + sc.prod_l, *pu = ix + sc.prod_u;
        const IT* lvl = ix + sc.lvl_lu;
        for (int lv = 0; lv < sc.n_levels; lv++) {
            int e0 = lvl[lv], e1 = lvl[lv + 1];
            if (fac) {
                for (int e = e0 + lane_e; e < e1; e += stride_e) {
                    int q0 = pp[e], q1 = pp[e + 1];
                    IT pv = pv_[e];
                    if (q1 == q0 && pv == NONE) continue;  // plain copy of W: already in place
                    // all operands are independent loads: issue them before the arithmetic
                    double v = lu[(long long)e * B];
                    double piv = pv != NONE ? lu[(long long)(int)pv * B] : 1.0;
                    int q = q0;
                    for (; q + 1 < q1; q += 2) {
                        double a0 = lu[(long long)(int)pl[q] * B], b0 = lu[(long long)(int)pu[q] * B];
                        double a1 = lu[(long long)(int)pl[q + 1] * B], b1 = lu[(long long)(int)pu[q + 1] * B];
                        v -= a0 * b0;
                        v -= a1 * b1;
                    }
                    if (q < q1) v -= lu[(long long)(int)pl[q] * B] * lu[(long long)(int)pu[q] * B];
                    if (pv != NONE) v /= piv;
                    lu[(long long)e * B] = v;
                }
            }
            __syncthreads();
        }
    }
    // 3. right-hand side, forward and backward substitution
    const IT* perm = ix + sc.perm;
    for (int i = lane_e; i < n.n_reduced; i += stride_e)
        xs_s[i * TILE + j] = m >= 0 ? rbs_solve_rhs<true>(n, nullptr, (int)perm[i], m) : 0.0;
    __syncthreads();
    {
        const IT *lvf = ix + sc.lvl_fwd, *lvb = ix + sc.lvl_bwd;
        const IT *fr = ix + sc.fwd_row, *fp = ix + sc.fwd_ptr, *fl = ix + sc.fwd_l, *fk = ix + sc.fwd_k;
        for (int lv = 1; lv < sc.n_fwd; lv++) {  // level 0 rows have no dependencies
            int p0 = lvf[lv], p1 = lvf[lv + 1];
            if (m >= 0) {
                for (int pos = p0 + lane_e; pos < p1; pos += stride_e) {
                    int i = fr[pos];
                    int q0 = fp[pos], q1 = fp[pos + 1];
                    double v = xs_s[i * TILE + j];
                    int q = q0;
                    for (; q + 1 < q1; q += 2) {
                        double a0 = lu[(long long)(int)fl[q] * B], a1 = lu[(long long)(int)fl[q + 1] * B];
                        v -= a0 * xs_s[(int)fk[q] * TILE + j];
                        v -= a1 * xs_s[(int)fk[q + 1] * TILE + j];
                    }
                    if (q < q1) v -= lu[(long long)(int)fl[q] * B] * xs_s[(int)fk[q] * TILE + j];
                    xs_s[i * TILE + j] = v;
                }
            }
            __syncthreads();
        }
        const IT *br = ix + sc.bwd_row, *bp = ix + sc.bwd_ptr, *bu = ix + sc.bwd_u, *bk = ix + sc.bwd_k;
        const IT* dg = ix + sc.diag;
        for (int lv = 0; lv < sc.n_bwd; lv++) {
            int p0 = lvb[lv], p1 = lvb[lv + 1];
            if (m >= 0) {
                for (int pos = p0 + lane_e; pos < p1; pos += stride_e) {
                    int i = br[pos];
                    int q0 = bp[pos], q1 = bp[pos + 1];
                    double d = lu[(long long)(int)dg[i] * B];
                    double v = xs_s[i * TILE + j];
                    int q = q0;
                    for (; q + 1 < q1; q += 2) {
                        double a0 = lu[(long long)(int)bu[q] * B], a1 = lu[(long long)(int)bu[q + 1] * B];
                        v -= a0 * xs_s[(int)bk[q] * TILE + j];
                        v -= a1 * xs_s[(int)bk[q + 1] * TILE + j];
                    }
                    if (q < q1) v -= lu[(long long)(int)bu[q] * B] * xs_s[(int)bk[q] * TILE + j];
                    xs_s[i * TILE + j] = v / d;
                }
            }
            __syncthreads();
        }
    }
    if (m >= 0)
        for (int i = lane_e; i < n.n_reduced; i += stride_e)
            x[(long long)(int)perm[i] * B + m] = xs_s[i * TILE + j];
}

// ---- step path, small and medium networks: factors in shared memory, LDU schedule ----------------
// TILE = 8 members per block.  L/U factors ([n_lu][8] doubles), solution vectors, the schedule and
// the assembly lists (which J entries land on which slot, which states reduce into which row) are
// all resident in shared memory, so every dependent level costs shared-memory latency only.
// The schedule is the "LDU" variant of ribasim_b200/symbolic.py: L is kept unscaled and the
// diagonal slots hold the pivot reciprocals r_k, so a slot depends only on pivots of earlier
// elimination rounds — one level per round and no division on the critical path:
//   slot(i,j) -= (slot(i,k) * r_k) * slot(k,j);   diagonal: r_i = 1 / slot(i,i)
//   forward   yhat_i = r_i * (b_i - sum_k L~[i,k] yhat_k)
//   backward  x_i = yhat_i - r_i * sum_j U[i,j] x_j
// Global traffic is the compulsory one: J values and du/z in, c out (plus the factors once when
// frozen-Newton iterations will reuse them).
struct LduSched {          // offsets (in index elements) into the packed schedule blob
    int prod_ptr, flag, prod_l, prod_u, prod_p, diag, perm;
    int fwd_row, fwd_ptr, fwd_l, fwd_k, bwd_row, bwd_ptr, bwd_u, bwd_k;
    int lvl, lvl_fwd, lvl_bwd;
    int wide, wide_fwd, wide_bwd;  // per level: 1 = split each product list over 4 lanes
    int total;
    int n_levels, n_fwd, n_bwd;
};

// Input of the shared-memory kernel, produced at full occupancy: the initial value of every
// factor slot (W = A J - I/h scattered into the LU pattern; ordered signed sums, code < 0 is the
// "- 1/h" of a diagonal entry) for members that refresh their Jacobian, and the reduced
// right-hand side b_r = A*((h du - z)/h) for every member in the Newton phase.
// one item (factor slot or reduced right-hand-side row) of one member: the arithmetic of
// k_assemble_linear, used by the cooperative window kernel
__device__ __forceinline__ void rbs_assemble_item(const Net& n, const int* __restrict__ dsrc_ptr,
                                                  const int* __restrict__ dsrc_code, int e, int m) {
    if (e < n.n_lu) {
        if (n.mjac[m] == 0) return;
        const double invh = 1.0 / n.mh[m];
        double v = 0.0;
        for (int k = dsrc_ptr[e]; k < dsrc_ptr[e + 1]; k++) {
            int code = dsrc_code[k];
            if (code < 0) { v = v - invh; continue; }
            double x = n.jv[(long long)(code >> 1) * n.B + m];
            v = (code & 1) ? v - x : v + x;
        }
        n.lu[(long long)e * n.B + m] = v;
    } else {
        int r = e - n.n_lu;
        n.br[(long long)r * n.B + m] = rbs_solve_rhs<true>(n, nullptr, r, m);
    }
}

// Two members per thread (positions j and j + 32 of a 64-member tile of the launch domain): the
// index loads of an item are shared by both, every load instruction is still a coalesced 256 B
// request.  The kernel is bound by instruction issue, not by bytes.
#define RBS_MPT 2
__global__ void k_assemble_linear(Net n, const int* __restrict__ dsrc_ptr,
                                  const int* __restrict__ dsrc_code) {
    RBS_PDL();
    const int n_tiles = (n.NA + 63) >> 6;
    long long tid = RBS_TID;
    const int lane = (int)(tid & 31);
    long long wid = tid >> 5;                 // (item, tile)
    int e = (int)(wid / n_tiles);
    if (e >= n.n_lu + n.n_reduced) return;
    int tile = (int)(wid - (long long)e * n_tiles);
    int j0 = tile * 64 + lane, j1 = j0 + 32;
    int m0 = j0 < n.NA ? RBS_MEMBER(n, j0) : -1, m1 = j1 < n.NA ? RBS_MEMBER(n, j1) : -1;
    if (m0 >= 0 && n.mphase[m0] != PHASE_NEWTON) m0 = -1;
    if (m1 >= 0 && n.mphase[m1] != PHASE_NEWTON) m1 = -1;
    if (e < n.n_lu) {
        if (m0 >= 0 && n.mjac[m0] == 0) m0 = -1;
        if (m1 >= 0 && n.mjac[m1] == 0) m1 = -1;
        if (m0 < 0 && m1 < 0) return;
        const int a0 = m0 >= 0 ? m0 : m1, a1 = m1 >= 0 ? m1 : m0;  // valid addresses for both
        const double invh0 = 1.0 / n.mh[a0], invh1 = 1.0 / n.mh[a1];
        double v0 = 0.0, v1 = 0.0;
        for (int k = dsrc_ptr[e]; k < dsrc_ptr[e + 1]; k++) {
            int code = dsrc_code[k];
            if (code < 0) { v0 = v0 - invh0; v1 = v1 - invh1; continue; }
            const double* row = n.jv + (long long)(code >> 1) * n.B;
            double x0 = row[a0], x1 = row[a1];
            if (code & 1) { v0 = v0 - x0; v1 = v1 - x1; } else { v0 = v0 + x0; v1 = v1 + x1; }
        }
        if (m0 >= 0) n.lu[(long long)e * n.B + m0] = v0;
        if (m1 >= 0) n.lu[(long long)e * n.B + m1] = v1;
    } else {
        int r = e - n.n_lu;
        double* dst = n.lane_iperm ? n.lu + (long long)(n.n_lu + n.lane_iperm[r]) * n.B : n.br + (long long)r * n.B;
        if (m0 >= 0) dst[m0] = rbs_solve_rhs<true>(n, nullptr, r, m0);
        if (m1 >= 0) dst[m1] = rbs_solve_rhs<true>(n, nullptr, r, m1);
    }
}

// phase clocks of block 0 (tuning aid, compiled in with -DRBS_LIN_PROFILE; see rbs_time_linear)
#ifdef RBS_LIN_PROFILE
__device__ long long g_lin_clk[8];
__device__ long long g_lin_lvl[64];
#define LIN_CLK(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_lin_clk[k] = clock64(); } while (0)
#else
#define LIN_CLK(k) do { } while (0)
#endif

template <class IT>
__global__ void __launch_bounds__(1024)
k_newton_linear_smem(Net n, LduSched sc, const IT* __restrict__ blob, int write_lu,
                     double* __restrict__ x) {
    RBS_PDL();
    constexpr int TILE = 8;
    extern __shared__ double smem[];
    double* lu_s = smem;                               // [n_lu][TILE]
    double* xs_s = lu_s + (size_t)n.n_lu * TILE;       // [n_reduced][TILE]
    IT* ix = (IT*)(xs_s + (size_t)n.n_reduced * TILE);
    __shared__ int mem_s[TILE], fac_s[TILE];
    __shared__ int any_fac;
    int m0 = blockIdx.x * TILE;
    int tm = min(TILE, n.NA - m0);
    if (tm <= 0) return;
    LIN_CLK(0);
    if (threadIdx.x == 0) any_fac = 0;
    __syncthreads();
    if (threadIdx.x < TILE) {
        int jj = threadIdx.x;
        int mm = jj < tm ? RBS_MEMBER(n, m0 + jj) : -1;
        if (mm >= 0 && n.mphase[mm] != PHASE_NEWTON) mm = -1;
        mem_s[jj] = mm;
        int f = mm >= 0 && n.mjac[mm] != 0;
        fac_s[jj] = f;
        if (f) any_fac = 1;
    }
    for (int w = threadIdx.x; w < sc.total; w += blockDim.x) ix[w] = blob[w];
    __syncthreads();
    LIN_CLK(1);
    const int j = threadIdx.x & (TILE - 1);
    const int slot = threadIdx.x >> 3, n_slot = blockDim.x >> 3;
    // in the level loops a warp owns one slot: 8 members x 4 lanes that share the slot's product
    // list (partial sums are combined with a fixed shuffle tree, so results are reproducible)
    const int g = (threadIdx.x >> 3) & 3;
    const int wslot = threadIdx.x >> 5, n_wslot = blockDim.x >> 5;
    const int m = mem_s[j];
    const bool fac = fac_s[j] != 0;
    const long long B = n.B;
    // 1. factor slots (initial values from k_assemble_linear, or the stored factors)
    if (m >= 0) {
        const double* lug = n.lu + m;
#pragma unroll 8
        for (int e = slot; e < n.n_lu; e += n_slot) lu_s[e * TILE + j] = lug[(long long)e * B];
    }
    __syncthreads();
    LIN_CLK(2);
    // 2. elimination: one level per round
    if (any_fac) {
        const IT *pp = ix + sc.prod_ptr, *fl_ = ix + sc.flag, *pl = ix + sc.prod_l, *pu = ix + sc.prod_u,
                 *ppv = ix + sc.prod_p;
        const IT *lvl = ix + sc.lvl, *wd = ix + sc.wide;
        for (int lv = 0; lv < sc.n_levels; lv++) {
            int e0 = lvl[lv], e1 = lvl[lv + 1];
            if (wd[lv]) {
                for (int e = e0 + wslot; e < e1; e += n_wslot) {  // warp-uniform
                    int q0 = pp[e], q1 = pp[e + 1];
                    bool isd = fl_[e] != 0;
                    if (q1 == q0 && !isd) continue;
                    double s = 0.0;
                    for (int q = q0 + g; q < q1; q += 4)
                        s += (lu_s[(int)pl[q] * TILE + j] * lu_s[(int)ppv[q] * TILE + j]) *
                             lu_s[(int)pu[q] * TILE + j];
                    s += __shfl_xor_sync(0xffffffffu, s, 8);
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    if (g == 0 && fac) {
                        double v = lu_s[e * TILE + j] - s;
                        lu_s[e * TILE + j] = isd ? 1.0 / v : v;
                    }
                }
            } else if (fac) {
                for (int e = e0 + slot; e < e1; e += n_slot) {
                    int q0 = pp[e], q1 = pp[e + 1];
                    bool isd = fl_[e] != 0;
                    if (q1 == q0 && !isd) continue;
                    double s = 0.0;
                    for (int q = q0; q < q1; q++)
                        s += (lu_s[(int)pl[q] * TILE + j] * lu_s[(int)ppv[q] * TILE + j]) *
                             lu_s[(int)pu[q] * TILE + j];
                    double v = lu_s[e * TILE + j] - s;
                    lu_s[e * TILE + j] = isd ? 1.0 / v : v;
                }
            }
            __syncthreads();
        }
        // keep the factors for the iterations that reuse them
        if (write_lu && fac)
            for (int e = slot; e < n.n_lu; e += n_slot) n.lu[(long long)e * B + m] = lu_s[e * TILE + j];
    }
    LIN_CLK(3);
    // 3. right-hand side in elimination order
    const IT* perm = ix + sc.perm;
    if (m >= 0) {
        const double* br = n.br + m;
#pragma unroll 4
        for (int i = slot; i < n.n_reduced; i += n_slot) xs_s[i * TILE + j] = br[(long long)(int)perm[i] * B];
    }
    __syncthreads();
    LIN_CLK(4);
    // 4. forward and backward substitution
    {
        const IT *lvf = ix + sc.lvl_fwd, *lvb = ix + sc.lvl_bwd, *dg = ix + sc.diag;
        const IT *fr = ix + sc.fwd_row, *fp = ix + sc.fwd_ptr, *fl = ix + sc.fwd_l, *fk = ix + sc.fwd_k;
        const IT *wdf = ix + sc.wide_fwd, *wdb = ix + sc.wide_bwd;
        for (int lv = 0; lv < sc.n_fwd; lv++) {
            int p0 = lvf[lv], p1 = lvf[lv + 1];
            if (wdf[lv]) {
                for (int pos = p0 + wslot; pos < p1; pos += n_wslot) {  // warp-uniform
                    int i = fr[pos];
                    double s = 0.0;
                    for (int q = fp[pos] + g; q < (int)fp[pos + 1]; q += 4)
                        s += lu_s[(int)fl[q] * TILE + j] * xs_s[(int)fk[q] * TILE + j];
                    s += __shfl_xor_sync(0xffffffffu, s, 8);
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    if (g == 0) xs_s[i * TILE + j] = (xs_s[i * TILE + j] - s) * lu_s[(int)dg[i] * TILE + j];
                }
            } else {
                for (int pos = p0 + slot; pos < p1; pos += n_slot) {
                    int i = fr[pos];
                    double s = 0.0;
                    for (int q = fp[pos]; q < (int)fp[pos + 1]; q++)
                        s += lu_s[(int)fl[q] * TILE + j] * xs_s[(int)fk[q] * TILE + j];
                    xs_s[i * TILE + j] = (xs_s[i * TILE + j] - s) * lu_s[(int)dg[i] * TILE + j];
                }
            }
            __syncthreads();
        }
        LIN_CLK(5);
        const IT *br = ix + sc.bwd_row, *bp = ix + sc.bwd_ptr, *bu = ix + sc.bwd_u, *bk = ix + sc.bwd_k;
        for (int lv = 1; lv < sc.n_bwd; lv++) {  // level 0 rows have no U entries: x = yhat
            int p0 = lvb[lv], p1 = lvb[lv + 1];
            if (wdb[lv]) {
                for (int pos = p0 + wslot; pos < p1; pos += n_wslot) {
                    int i = br[pos];
                    double s = 0.0;
                    for (int q = bp[pos] + g; q < (int)bp[pos + 1]; q += 4)
                        s += lu_s[(int)bu[q] * TILE + j] * xs_s[(int)bk[q] * TILE + j];
                    s += __shfl_xor_sync(0xffffffffu, s, 8);
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    if (g == 0) xs_s[i * TILE + j] = xs_s[i * TILE + j] - lu_s[(int)dg[i] * TILE + j] * s;
                }
            } else {
                for (int pos = p0 + slot; pos < p1; pos += n_slot) {
                    int i = br[pos];
                    double s = 0.0;
                    for (int q = bp[pos]; q < (int)bp[pos + 1]; q++)
                        s += lu_s[(int)bu[q] * TILE + j] * xs_s[(int)bk[q] * TILE + j];
                    xs_s[i * TILE + j] = xs_s[i * TILE + j] - lu_s[(int)dg[i] * TILE + j] * s;
                }
            }
            __syncthreads();
        }
    }
    LIN_CLK(6);
    if (m >= 0)
        for (int i = slot; i < n.n_reduced; i += n_slot)
            x[(long long)(int)perm[i] * B + m] = xs_s[i * TILE + j];
    LIN_CLK(7);
}

// ---- entry schedule: factorisation with fused forward substitution + backward substitution -------
// See ribasim_b200/symbolic.py (build_entries).  A block keeps the slots of a tile of TILE members
// in shared memory as S[slot][TILE].  TILE/2 threads own one entry of a level, each thread two
// members (one 128-bit shared-memory access per operand), and accumulate the entry's products
//     acc = sum_q (S[l_q] * S[p_q]) * S[u_q];  v = S[e] (* S[r] if SCALE) - acc;  S[e] = RCP ? 1/v : v
// In levels with few entries and long lists a team of 2^team_log lanes shares an entry (products
// strided over the team, partial sums combined by a fixed shuffle butterfly: reproducible).
// Levels only read slots of earlier levels and are separated by block barriers.
//
// The level loop is a chain of dependent shared-memory loads and fp64 operations: a block on its
// own keeps an SM's schedulers ~20% busy.  So the tile is small (TILE = 2: 46 KB of slots for the
// 500-basin network) and the schedule is read through L1 from global memory instead of being
// copied into every block's shared memory: four or more blocks share an SM and fill each other's
// barrier and latency gaps.
// Members that reuse stored factors (`fac` bit clear) only update the right-hand-side column.
#define RBS_ENT_RCP 0x100
#define RBS_ENT_SCALE 0x200
#define RBS_ENT_FRESH 0x400
#define RBS_ENT_LANES 256
// The schedule as the device reads it: one 64-bit word per level, entry and product, so that a
// lane fetches its work with one load each (the level loop is bound by instruction count and
// dependent-load latency, not by bytes).
//   level word   : e0 | e1 << 16 | team_log << 32
//   entry word   : target slot | q0 << 16 | (count | RCP | SCALE) << 32
//   product word : l | p << 16 | u << 32
struct EntSched {
    const unsigned long long *lvl, *ent, *prod;   // lvl, ent, prod are contiguous (lvl first)
    const uint16_t* perm;
    int n_words;              // 64-bit words of lvl + ent + prod
    int smem_sched;           // copy the schedule into shared memory (after the slots)
    int n_elim, n_bwd, n_slots, y0;
    // compact form (symbolic.py `_compact_slots`): fill-ins take over the slots of dead entries, so
    // the tile needs peak-live instead of all slots.  `load[k]` = LU entry | slot << 16 of the
    // entries that start from the assembled values; a FRESH entry ignores what its slot held.
    // The factors are gone afterwards: full-Newton path only.  n_load = 0: LU entry e in slot e.
    const unsigned* load;
    int n_load;
};

__device__ __forceinline__ void rbs_cp_async8(void* smem_dst, const void* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void rbs_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// facmask: bit 0 / 1 set = first / second member of this thread's pair refreshes its factors
// schedule word: through the read-only global path when the schedule lives in global memory (a
// generic load would decide the address space per access), a plain load when it is in shared memory
template <bool GL> __device__ __forceinline__ unsigned long long rbs_ldw(const unsigned long long* p) {
    return GL ? __ldg(p) : *p;
}
template <int TILE, bool GL>
__device__ __forceinline__ void rbs_ent_levels_impl(double* S, const EntSched& es, int lv0, int lv1,
                                                    unsigned facmask) {
    constexpr int TPL = TILE / 2;            // threads per entry lane
    constexpr int EL = RBS_ENT_LANES;
    const int lane_e = threadIdx.x / TPL;
    const int warp_lane0 = (int)(threadIdx.x >> 5) * (32 / TPL);
    const double* Sc = S + (threadIdx.x % TPL) * 2;  // this thread's member pair
    const unsigned long long* ent = es.ent;
    const unsigned long long* prod = es.prod;
    for (int lv = lv0; lv < lv1; lv++) {
        const unsigned long long lw = rbs_ldw<GL>(es.lvl + lv);
        const int e0 = (int)(lw & 0xFFFFu), e1 = (int)((lw >> 16) & 0xFFFFu), tl = (int)(lw >> 32);
        // warps whose lanes get no entry in the first pass have none in the later ones either:
        // they go straight to the barrier
        if (warp_lane0 < min((e1 - e0) << tl, EL)) {
            const int ts = 1 << tl, t = lane_e & (ts - 1), stride = EL >> tl;
            for (int base = e0; base < e1; base += stride) {  // warp-uniform trip count
                const int k = base + (lane_e >> tl);
                const bool valid = k < e1;
                double a0 = 0.0, a1 = 0.0;
                unsigned e = 0, q0 = 0, cf = 0;
                if (valid) {
                    const unsigned long long ew = rbs_ldw<GL>(ent + k);
                    e = (unsigned)ew & 0xFFFFu; q0 = ((unsigned)ew) >> 16; cf = (unsigned)(ew >> 32);
                    const unsigned q1 = q0 + (cf & 0xFFu);
#pragma unroll 2
                    for (unsigned q = q0 + t; q < q1; q += ts) {
                        const unsigned long long pw = rbs_ldw<GL>(prod + q);
                        const double2 l = *(const double2*)(Sc + ((unsigned)pw & 0xFFFFu) * TILE);
                        const double2 p = *(const double2*)(Sc + (((unsigned)pw) >> 16) * TILE);
                        const double2 u = *(const double2*)(Sc + ((unsigned)(pw >> 32) & 0xFFFFu) * TILE);
                        a0 += (l.x * p.x) * u.x;
                        a1 += (l.y * p.y) * u.y;
                    }
                }
                if (tl) {  // uniform; teams are aligned groups of lanes
                    a0 += __shfl_xor_sync(0xffffffffu, a0, TPL);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, TPL);
                    if (tl > 1) {
                        a0 += __shfl_xor_sync(0xffffffffu, a0, TPL * 2);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, TPL * 2);
                        if (tl > 2) {
                            a0 += __shfl_xor_sync(0xffffffffu, a0, TPL * 4);
                            a1 += __shfl_xor_sync(0xffffffffu, a1, TPL * 4);
                        }
                    }
                }
                if (valid && t == 0) {
                    double2* T = (double2*)(const_cast<double*>(Sc) + e * TILE);
                    double2 o = *T;
                    if (cf & RBS_ENT_FRESH) { o.x = 0.0; o.y = 0.0; }
                    double2 v = o;
                    if (cf & RBS_ENT_SCALE) {
                        // the reciprocal pivot: p of the first product, or q0 itself without products
                        const unsigned r = (cf & 0xFFu) ? (((unsigned)rbs_ldw<GL>(prod + q0)) >> 16) : q0;
                        const double2 rv = *(const double2*)(Sc + r * TILE);
                        v.x *= rv.x; v.y *= rv.y;
                    }
                    v.x -= a0; v.y -= a1;
                    if (cf & RBS_ENT_RCP) { v.x = 1.0 / v.x; v.y = 1.0 / v.y; }
                    if ((int)e < es.y0 && facmask != 3u) {
                        if (!(facmask & 1u)) v.x = o.x;
                        if (!(facmask & 2u)) v.y = o.y;
                    }
                    *T = v;
                }
            }
        }
        __syncthreads();
#ifdef RBS_LVL_PROFILE
        if (blockIdx.x == 0 && threadIdx.x == 0 && lv < 64) g_lin_lvl[lv] = clock64();
#endif
    }
}

template <int TILE>
__device__ __forceinline__ void rbs_ent_levels(double* S, const EntSched& es, int lv0, int lv1,
                                               unsigned facmask) {
    if (es.smem_sched == 0) rbs_ent_levels_impl<TILE, true>(S, es, lv0, lv1, facmask);
    else rbs_ent_levels_impl<TILE, false>(S, es, lv0, lv1, facmask);
}

// Newton linear solve of one member tile per block on the entry schedule; input as for
// k_newton_linear_smem: the initial slot values / stored factors in `lu`, the reduced right-hand
// side in `br` (k_assemble_linear).
// the solve of the member tile [m0, m0 + TILE) of the launch domain by one block (block barriers
// inside; `S` = the block's dynamic shared memory)
template <int TILE>
__device__ __forceinline__ void rbs_lin_tile(const Net& n, EntSched es, double* S, int m0, int write_lu,
                                             double* x) {
    constexpr int NT = RBS_ENT_LANES * TILE / 2;
    const uint16_t* __restrict__ perm = es.perm;
    __shared__ int mem_s[TILE], fac_s[TILE];
    int tm = min(TILE, n.NA - m0);
    if (tm <= 0) return;
    LIN_CLK(0);
    if (threadIdx.x < TILE) {
        int jj = threadIdx.x;
        int mm = jj < tm ? RBS_MEMBER(n, m0 + jj) : -1;
        if (mm >= 0 && n.mphase[mm] != PHASE_NEWTON) mm = -1;
        mem_s[jj] = mm;
        fac_s[jj] = mm >= 0 && n.mjac[mm] != 0;
    }
    if (es.smem_sched == 3) {
        // 3: the product words only — they are the gathered loads of the level loop (a lane's
        // products lie anywhere in the list); level and entry words are coalesced and stay behind L1
        unsigned long long* sw = (unsigned long long*)(S + (size_t)es.n_slots * TILE);
        const int n_prod = es.n_words - (int)(es.prod - es.lvl);
        for (int w = threadIdx.x; w < n_prod; w += NT) sw[w] = es.prod[w];
        es.prod = sw;
    } else if (es.smem_sched) {
        // 1: the whole schedule behind the slots; 2: level and entry words only (a lane's entry
        // word heads its dependent chain in every level; the product words stay behind L1)
        unsigned long long* sw = (unsigned long long*)(S + (size_t)es.n_slots * TILE);
        const int n_copy = es.smem_sched == 2 ? (int)(es.prod - es.lvl) : es.n_words;
        for (int w = threadIdx.x; w < n_copy; w += NT) sw[w] = es.lvl[w];
        es.ent = sw + (es.ent - es.lvl);
        if (es.smem_sched != 2) es.prod = sw + (es.prod - es.lvl);
        es.lvl = sw;
    }
    __syncthreads();
    LIN_CLK(1);
    const int j = threadIdx.x % TILE;
    const int slot = threadIdx.x / TILE, n_slot = NT / TILE;
    const int m = mem_s[j];
    const long long B = n.B;
    if (m >= 0) {
        const double* lug = n.lu + m;
        if (es.n_load == 0) {
            for (int e = slot; e < n.n_lu; e += n_slot) rbs_cp_async8(S + e * TILE + j, lug + (long long)e * B);
        } else {
            for (int k = slot; k < es.n_load; k += n_slot) {
                const unsigned w = es.load[k];
                rbs_cp_async8(S + (w >> 16) * TILE + j, lug + (long long)(w & 0xFFFFu) * B);
            }
        }
        const double* br = n.br + m;
        for (int i = slot; i < n.n_reduced; i += n_slot)
            rbs_cp_async8(S + (es.y0 + i) * TILE + j, br + (long long)(int)perm[i] * B);
    }
    rbs_cp_async_wait_all();
    __syncthreads();
    LIN_CLK(2);
    const int c2 = (threadIdx.x % (TILE / 2)) * 2;
    const unsigned facmask = (fac_s[c2] ? 1u : 0u) | (fac_s[c2 + 1] ? 2u : 0u);
    rbs_ent_levels<TILE>(S, es, 0, es.n_elim, facmask);
    LIN_CLK(3);
    LIN_CLK(4);
    LIN_CLK(5);
    rbs_ent_levels<TILE>(S, es, es.n_elim, es.n_elim + es.n_bwd, facmask);
    LIN_CLK(6);
    if (m >= 0) {
        for (int i = slot; i < n.n_reduced; i += n_slot)
            x[(long long)(int)perm[i] * B + m] = S[(es.y0 + i) * TILE + j];
        if (write_lu && fac_s[j] && es.n_load == 0)
            for (int e = slot; e < n.n_lu; e += n_slot) n.lu[(long long)e * B + m] = S[e * TILE + j];
    }
    LIN_CLK(7);
}
template <int TILE>
__global__ void __launch_bounds__(RBS_ENT_LANES * TILE / 2)
k_newton_linear_ent(Net n, EntSched es, int write_lu, double* __restrict__ x) {
    RBS_PDL();
    extern __shared__ double smem[];
    rbs_lin_tile<TILE>(n, es, smem, blockIdx.x * TILE, write_lu, x);
}

// ---- limit_flow! (solve.jl:840-984); needs storage/level at the proposed u ------------------------
__device__ __forceinline__ double rbs_min_low_storage_factor(const Net& n, int kind, int idx, int m) {
    if (kind != KIND_BASIN) return 1.0;
    long long o = (long long)idx * n.B + m;
    double thr = n.low_storage_threshold[idx];
    return rbs_reduction_factor(fmin(n.storage[o], n.storage_prev[o]) - 2 * thr, thr);
}
// flow-rate bounds of state s over the last step; false = unbounded (Manning, integral)
__device__ __forceinline__ bool rbs_flow_bounds(const Net& n, int s, int m,
                                                const int* __restrict__ from_ts,
                                                const double* __restrict__ alloc_total, double& lo,
                                                double& hi) {
    if (s < n.n_conn) {
        int ct = n.conn_type[s], k = n.conn_idx[s];
        if (ct == CT_TRC) { lo = 0.0; hi = RBS_INF; return true; }
        if (ct == CT_PUMP) {
            int row = rbs_dc_row(n, n.pump_dc, n.pump_cs_ptr, k, m);
            lo = RBS_P(row, n.pump_cs_min_flow_rate, n.pump_min_flow_rate, k);
            hi = RBS_P(row, n.pump_cs_max_flow_rate, n.pump_max_flow_rate, k);
            return true;
        }
        if (ct == CT_OUTLET) {
            int row = rbs_dc_row(n, n.outlet_dc, n.outlet_cs_ptr, k, m);
            lo = RBS_P(row, n.outlet_cs_min_flow_rate, n.outlet_min_flow_rate, k);
            hi = RBS_P(row, n.outlet_cs_max_flow_rate, n.outlet_max_flow_rate, k);
            return true;
        }
        if (ct == CT_LR) {
            int row = rbs_dc_row(n, n.lr_dc, n.lr_cs_ptr, k, m);
            hi = RBS_P(row, n.lr_cs_max_flow_rate, n.lr_max_flow_rate, k);
            lo = -hi;
            return true;
        }
        if (ct == CT_UD_IN) {
            int i = n.ud_of_link[k];
            if (from_ts[i]) { lo = 0.0; hi = RBS_INF; return true; }
            int n_links = n.ud_link_ptr[i + 1] - n.ud_link_ptr[i];
            double equal_split = n_links == 0 ? 0.0 : alloc_total[i] / n_links;
            double alloc = n.ud_link_alloc[k];
            double q_k_max = isinf(alloc) ? equal_split : alloc;
            double fb = rbs_min_low_storage_factor(n, n.src_kind[s], n.src_idx[s], m);
            double fl = 1.0;
            if (n.src_kind[s] == KIND_BASIN) {
                long long bo = (long long)n.src_idx[s] * n.B + m;
                fl = rbs_reduction_factor(fmin(n.level[bo], n.level_prev[bo]) - n.ud_min_level[i] -
                                              2 * n.ldt, n.ldt);
            }
            lo = fb * fl * q_k_max; hi = q_k_max;
            return true;
        }
        return false;  // UserDemand outflow, ManningResistance: not limited
    }
    if (s < n.off_infil) { lo = 0.0; hi = RBS_INF; return true; }
    if (s < n.off_integral) {
        int b = s - n.off_infil;
        double fmin_ = rbs_min_low_storage_factor(n, KIND_BASIN, b, m);
        double infil = rbs_fw(n, n.infiltration, n.w_infiltration, n.fw_sb,
                              n.fw_on ? n.mstep[m] : 0)[(long long)b * n.B + m];
        lo = fmin_ * infil; hi = infil;
        return true;
    }
    return false;
}
__device__ __forceinline__ double rbs_clamp_state(double v, double up, double lo, double hi, double dt) {
    double a = up + lo * dt, b = up + hi * dt;
    return v > b ? b : (v < a ? a : v);
}
// seam: clamp u in place against uprev over a shared dt
__global__ void k_limit_flow(Net n, double* __restrict__ u, const double* __restrict__ uprev,
                             double dt, const int* __restrict__ from_ts,
                             const double* __restrict__ alloc_total) {
    RBS_PDL();
    long long tid = RBS_TID;
    int s = (int)(tid / n.NA);
    if (s >= n.off_integral) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)s * n.NA));
    long long o = (long long)s * n.B + m;
    double lo, hi;
    if (!rbs_flow_bounds(n, s, m, from_ts, alloc_total, lo, hi)) return;
    u[o] = rbs_clamp_state(u[o], uprev[o], lo, hi, dt);
}

// ---- Newton update (dolinsolve tail + compute_step!/apply_step!) --------------------
// dz = gamma*h * (J_i c - b),  b = (h k - z)/(gamma h);  z <- z - dz;
// atmp = dz / (abstol + max(|uprev|, |uprev + z_old|) * reltol); partial sums of atmp^2.
#define RBS_DZ_ROWS 8
// Two members per thread (positions j and j + 32 of a 64-member tile): the row pointers and
// column indices of J_intermediate are loaded once for both.
// one (64-member tile, row block) unit; tx in [0, 32), ty in [0, RBS_DZ_ROWS); all threads of the
// block take part (block barrier inside)
__device__ __forceinline__ void rbs_update_unit(const Net& n, const double* c, double abstol,
                                                double reltol, int rows_per_block, int tile64, int part,
                                                int tx, int ty, double (*sh)[RBS_DZ_ROWS][33]) {
    const int j0 = tile64 * 64 + tx, j1 = j0 + 32;
    int m0 = j0 < n.NA ? RBS_MEMBER(n, j0) : -1, m1 = j1 < n.NA ? RBS_MEMBER(n, j1) : -1;
    if (m0 >= 0 && n.mphase[m0] != PHASE_NEWTON) m0 = -1;
    if (m1 >= 0 && n.mphase[m1] != PHASE_NEWTON) m1 = -1;
    const int r0 = part * rows_per_block;
    double acc0 = 0.0, acc1 = 0.0;
    if (m0 >= 0 || m1 >= 0) {
        const int a0 = m0 >= 0 ? m0 : m1, a1 = m1 >= 0 ? m1 : m0;  // valid addresses for both
        const double dt0 = n.mh[a0], dt1 = n.mh[a1], inv0 = 1.0 / dt0, inv1 = 1.0 / dt1;
        const int r1 = min(r0 + rows_per_block, n.n_state);
        for (int r = r0 + ty; r < r1; r += RBS_DZ_ROWS) {
            const long long o0 = (long long)r * n.B + a0, o1 = (long long)r * n.B + a1;
            double jc0 = 0.0, jc1 = 0.0;
            for (int e = n.jrow_ptr[r]; e < n.jrow_ptr[r + 1]; e++) {
                const double* jrow = n.jv + (long long)e * n.B;
                const double* crow = c + (long long)n.jcol[e] * n.B;
                jc0 += jrow[a0] * crow[a0];
                jc1 += jrow[a1] * crow[a1];
            }
            const double zo0 = n.z[o0], zo1 = n.z[o1];
            const double b0 = (dt0 * n.du[o0] - zo0) * inv0, b1 = (dt1 * n.du[o1] - zo1) * inv1;
            const double dz0 = (jc0 - b0) * dt0, dz1 = (jc1 - b1) * dt1;
            const double up0 = n.uprev[o0], up1 = n.uprev[o1];
            const double s0 = abstol + fmax(fabs(up0), fabs(up0 + zo0)) * reltol;
            const double s1 = abstol + fmax(fabs(up1), fabs(up1 + zo1)) * reltol;
            const double q0 = dz0 / s0, q1 = dz1 / s1;
            if (m0 >= 0) { acc0 += q0 * q0; n.z[o0] = zo0 - dz0; }
            if (m1 >= 0) { acc1 += q1 * q1; n.z[o1] = zo1 - dz1; }
        }
    }
    sh[0][ty][tx] = acc0;
    sh[1][ty][tx] = acc1;
    __syncthreads();
    if (ty < 2) {
        const int mm = ty == 0 ? m0 : m1;
        if (mm >= 0) {
            double s = 0.0;
            for (int y = 0; y < RBS_DZ_ROWS; y++) s += sh[ty][y][tx];
            n.norm_part[(long long)part * n.B + mm] = s;
        }
    }
}
// nlsolve!'s convergence test fused into the update: the last row-block of a 64-member tile to
// finish (a counter per tile) sums the partial norms in their fixed order and tests its members
struct ConvArgs {
    int* tile_counter;  // nullptr = no test here (the caller runs rbs_newton_converge itself)
    int n_part, max_iter, full_newton, early_exit;
    double kappa;
};
__device__ __forceinline__ void rbs_newton_converge(const Net& n, int m, int n_part, int max_iter,
                                                    double kappa, int full_newton, int early_exit);
__global__ void k_newton_update(Net n, const double* __restrict__ c, double abstol, double reltol,
                                int rows_per_block, ConvArgs cv) {
    RBS_PDL();
    __shared__ double sh[2][RBS_DZ_ROWS][33];
    __shared__ int last;
    rbs_update_unit(n, c, abstol, reltol, rows_per_block, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y, sh);
    if (!cv.tile_counter) return;
    const int t = threadIdx.y * 32 + threadIdx.x;
    __threadfence();
    __syncthreads();
    if (t == 0) last = atomicAdd(&cv.tile_counter[blockIdx.x], 1) == (int)gridDim.y - 1;
    __syncthreads();
    if (!last) return;
    if (t == 0) cv.tile_counter[blockIdx.x] = 0;
    __threadfence();
    if (t < 64) {
        const int j = blockIdx.x * 64 + t;
        if (j < n.NA) {
            const int m = RBS_MEMBER(n, j);
            if (n.mphase[m] == PHASE_NEWTON)
                rbs_newton_converge(n, m, cv.n_part, cv.max_iter, cv.kappa, cv.full_newton, cv.early_exit);
        }
    }
}

// per-member convergence bookkeeping of nlsolve! (OrdinaryDiffEqNonlinearSolve);
// a member whose iteration ends (converged or failed) moves to PHASE_LIMIT.
__device__ __forceinline__ void rbs_newton_converge(const Net& n, int m, int n_part, int max_iter,
                                                    double kappa, int full_newton, int early_exit) {
    int iter = n.miter[m];
    double s = 0.0;
#pragma unroll 8
    for (int p = 0; p < n_part; p++) s += __ldcg(&n.norm_part[(long long)p * n.B + m]);  // written by other blocks
    double ndz = sqrt(s / n.n_state);
    double eta = n.eta[m];
    int st = 0;  // 0 continue, 1 converged, 2 failed
    if (!isfinite(ndz)) st = 2;
    else {
        double theta = 0.0;
        if (iter > 1) {
            theta = ndz / n.ndz_prev[m];
            if (theta > 2.0) st = 2;
            // Hairer-Wanner: the remaining iterations cannot bring eta*ndz below kappa
            else if (early_exit && (theta >= 1.0 ||
                     ndz * pow(theta, (double)(max_iter - iter)) > kappa * (1.0 - theta))) st = 2;
            else eta = theta / (1.0 - theta);
        }
        if (st == 0 && ((iter == 1 && ndz < 1e-5) || (iter > 1 && eta >= 0.0 && eta * ndz < kappa)))
            st = 1;
    }
    if (st == 0 && iter >= max_iter) st = 2;
    n.ndz_prev[m] = ndz;
    n.ndz[m] = ndz;
    n.eta[m] = eta;
    n.cnt_iters[m] += 1;
    if (st == 0) {
        n.miter[m] = iter + 1;
        n.mjac[m] = full_newton;
    } else {
        n.mconv[m] = st == 1;
        n.mphase[m] = PHASE_LIMIT;
    }
}

// Ordered compaction of the members [m_lo, m_hi) that still have work: `alist_out` = members not
// DONE, `llist_out` = members in PHASE_LIMIT, counts[0..1] their lengths.  One block; order is
// kept so that a mostly-active ensemble stays coalesced.  Integer work only.
// With converge != 0 the members in the Newton phase first get the convergence test of the
// iteration they just finished (all of them iterated in this pass).
__global__ void k_compact(Net n, int m_lo, int m_hi, int* __restrict__ alist_out,
                          int* __restrict__ llist_out, int* __restrict__ counts, int converge,
                          int n_part, int max_iter, double kappa, int full_newton, int early_exit) {
    RBS_PDL();
    __shared__ int sa[1024], sl[1024];
    int t = threadIdx.x, T = blockDim.x;
    int nm = m_hi - m_lo;
    int per = (nm + T - 1) / T;
    int lo = m_lo + min(t * per, nm), hi = min(lo + per, m_hi);
    int ca = 0, cl = 0;
    for (int m = lo; m < hi; m++) {
        if (converge && n.mphase[m] == PHASE_NEWTON)
            rbs_newton_converge(n, m, n_part, max_iter, kappa, full_newton, early_exit);
        int ph = n.mphase[m];
        ca += ph != PHASE_DONE;
        cl += ph == PHASE_LIMIT;
    }
    sa[t] = ca; sl[t] = cl;
    __syncthreads();
    for (int off = 1; off < T; off <<= 1) {  // inclusive Hillis-Steele scan
        int va = t >= off ? sa[t - off] : 0, vl = t >= off ? sl[t - off] : 0;
        __syncthreads();
        sa[t] += va; sl[t] += vl;
        __syncthreads();
    }
    int pa = sa[t] - ca, pl = sl[t] - cl;
    for (int m = lo; m < hi; m++) {
        int ph = n.mphase[m];
        if (ph != PHASE_DONE) alist_out[pa++] = m;
        if (ph == PHASE_LIMIT) llist_out[pl++] = m;
    }
    if (t == T - 1) { counts[0] = sa[t]; counts[1] = sl[t]; }
    // counts[2]: outer steps of the window that every member of the group has completed (the
    // host streams the result slices of finished steps back while the window is still running)
    __syncthreads();
    int mn = 0x7fffffff;
    for (int m = lo; m < hi; m++) mn = min(mn, n.mstep[m]);
    sa[t] = mn;
    __syncthreads();
    for (int off = T >> 1; off > 0; off >>= 1) {
        if (t < off) sa[t] = min(sa[t], sa[t + off]);
        __syncthreads();
    }
    if (t == 0) counts[2] = sa[0];
}

// ---- sub-stepping state machine -------------------------------------------------------------------
// Exact cumulative forcing at the end of the member's current attempt (solve.jl:137-150, 86-99);
// on ACCEPT first advances the cumulatives by the accepted length (callback.jl:125-182).
__device__ __forceinline__ void rbs_exact_update(const Net& n, int b, int m, bool accept, double h_acc,
                                                 bool next, double h_next, int ws_acc, int ws_next) {
    // ws_acc / ws_next: outer step (of a forced window) the accepted attempt belonged to / the
    // next attempt belongs to
    long long o = (long long)b * n.B + m;
    double cp = n.cum_precipitation[o], cr = n.cum_surface_runoff[o], cd = n.cum_drainage[o];
    if (accept) {
        double P = n.fixed_area[b] * rbs_fw(n, n.precipitation, n.w_precipitation, n.fw_sb, ws_acc)[o];
        double R = rbs_fw(n, n.surface_runoff, n.w_surface_runoff, n.fw_sb, ws_acc)[o];
        double D = rbs_fw(n, n.drainage, n.w_drainage, n.fw_sb, ws_acc)[o];
        cp += P * h_acc; cr += h_acc * R; cd += h_acc * D;
        n.cum_precipitation[o] = cp; n.cum_surface_runoff[o] = cr; n.cum_drainage[o] = cd;
    }
    double P = n.fixed_area[b] * rbs_fw(n, n.precipitation, n.w_precipitation, n.fw_sb, ws_next)[o];
    double R = rbs_fw(n, n.surface_runoff, n.w_surface_runoff, n.fw_sb, ws_next)[o];
    double D = rbs_fw(n, n.drainage, n.w_drainage, n.fw_sb, ws_next)[o];
    double ex = ((cp + P * h_next) + (cr + h_next * R)) + (cd + h_next * D);
    const double* fr_acc = rbs_fw(n, n.fb_rate, n.w_fb_rate, n.fw_sf, ws_acc);
    const double* fr_next = rbs_fw(n, n.fb_rate, n.w_fb_rate, n.fw_sf, ws_next);
    for (int k = n.bfb_ptr[b]; k < n.bfb_ptr[b + 1]; k++) {
        long long f = (long long)n.bfb_idx[k] * n.B + m;
        double fc = n.fb_cum[f];
        if (accept) { fc += fr_acc[f] * h_acc; n.fb_cum[f] = fc; }
        ex += fc + fr_next[f] * h_next;
    }
    if (next) n.exact[o] = ex;
}

// start of an outer step [t0, t0 + dt]: uprev <- u, first attempt length, predictor, exact
__global__ void k_step_begin(Net n, double dt, int predictor) {
    RBS_PDL();
    long long tid = RBS_TID;
    long long total = (long long)n.n_state * n.B;
    int m = (int)(tid % n.B);
    double hnom = fmin(n.mhnom[m] > 0.0 ? n.mhnom[m] : dt, dt);
    double h = hnom;  // elapsed = 0: the remainder is dt
    if (tid < total) {
        double hl = n.mhlast[m];
        const double u0 = n.u[tid];
        n.uprev[tid] = u0;
        if (n.fw_on && n.w_flow) n.ustep[tid] = u0;
        n.z[tid] = (predictor && hl > 0.0) ? n.zlast[tid] * (h / hl) : 0.0;
    }
    long long bt = tid - total;
    if (bt >= 0 && bt < (long long)n.n_basin * n.B)
        rbs_exact_update(n, (int)(bt / n.B), m, false, 0.0, true, h, 0, 0);
    long long mt = bt - (long long)n.n_basin * n.B;
    if (mt >= 0 && mt < n.B) {
        n.mhnom[m] = hnom;
        n.mh[m] = h;
        n.mlast[m] = 1 * (hnom >= dt);
        n.melapsed[m] = 0.0;
        n.mphase[m] = PHASE_NEWTON;
        n.miter[m] = 1;
        n.mjac[m] = 1;
        n.mneg[m] = 0;
        n.mneg2[m] = 0;
        n.mclamp[m] = 0;
        n.maction[m] = ACTION_NONE;
        n.status[m] = 0;
        n.mstep[m] = 0;
        n.eta[m] = pow(fmax(n.eta[m], 2.220446049250313e-16), 0.8);
        n.ndz_prev[m] = 0.0;
    }
}

// PHASE_LIMIT members: u <- clamp(uprev + z) (limit_flow! on the proposed state)
__device__ __forceinline__ void rbs_step_limit_item(const Net& n, const int* __restrict__ from_ts,
                                                    const double* __restrict__ alloc_total, int s, int m) {
    if (n.mphase[m] != PHASE_LIMIT) return;
    long long o = (long long)s * n.B + m;
    double up = n.uprev[o];
    double v = up + n.z[o];
    double lo, hi;
    if (rbs_flow_bounds(n, s, m, from_ts, alloc_total, lo, hi)) {
        double vc = rbs_clamp_state(v, up, lo, hi, n.mh[m]);
        if (vc != v) n.mclamp[m] = 1;  // every writer stores the same value
        v = vc;
    }
    n.u[o] = v;
}

__global__ void k_step_limit(Net n, const int* __restrict__ from_ts,
                             const double* __restrict__ alloc_total) {
    RBS_PDL();
    long long tid = RBS_TID;
    int s = (int)(tid / n.NA);
    if (s >= n.n_state) return;
    int m = RBS_MEMBER(n, (int)(tid - (long long)s * n.NA));
    rbs_step_limit_item(n, from_ts, alloc_total, s, m);
}

// PHASE_LIMIT members: accept or reject the attempt, choose the next attempt
__device__ __forceinline__ void rbs_step_decide_item(const Net& n, double dt, double h_min, int max_halvings,
                                                     int grow_iters, int n_steps, int shrink_log2, int m) {
    if (n.mphase[m] != PHASE_LIMIT) { n.maction[m] = ACTION_NONE; return; }
    bool conv = n.mconv[m] != 0, last = n.mlast[m] != 0;
    bool neg = (n.mclamp[m] ? n.mneg2[m] : n.mneg[m]) != 0;
    n.mneg2[m] = 0;
    n.mclamp[m] = 0;
    double h = n.mh[m], hnom = n.mhnom[m], elapsed = n.melapsed[m];
    bool can_halve = max_halvings > 0 && 0.5 * h >= h_min * (1.0 - 1e-12);
    n.mneg[m] = 0;
    if ((conv && !neg) || !can_halve) {
        elapsed = last ? dt : elapsed + h;
        n.cnt_sub[m] += 1;
        if (!conv) n.status[m] |= 1;
        if (neg) n.status[m] |= 2;
        n.mhlast[m] = h;
        if (conv && n.miter[m] <= grow_iters) hnom = fmin(2.0 * hnom, dt);
        n.maction[m] = ACTION_ACCEPT;
    } else {
        hnom = fmax(ldexp(h, -shrink_log2), h_min);
        n.eta[m] = 1.0;
        n.cnt_rej[m] += 1;
        n.maction[m] = ACTION_REJECT;
    }
    n.mroll[m] = elapsed >= dt;
    if (elapsed >= dt) {
        // outer step finished: the member goes straight on to its next outer step of the
        // window (members never wait for each other), or is done
        int ks = n.mstep[m] + 1;
        n.mstep[m] = ks;
        if (ks >= n_steps) {
            n.melapsed[m] = elapsed;
            n.mhnom[m] = hnom;
            n.mphase[m] = PHASE_DONE;
            return;
        }
        elapsed = 0.0;
    }
    n.melapsed[m] = elapsed;
    n.mhnom[m] = hnom;
    double rem = dt - elapsed;
    bool nlast = hnom >= rem;
    n.mlast[m] = nlast;
    n.mh[m] = nlast ? rem : hnom;
    n.mphase[m] = PHASE_NEWTON;
    n.miter[m] = 1;
    n.mjac[m] = 1;
    n.eta[m] = pow(fmax(n.eta[m], 2.220446049250313e-16), 0.8);
    n.ndz_prev[m] = 0.0;
}

__global__ void k_step_decide(Net n, double dt, double h_min, int max_halvings, int grow_iters,
                              int n_steps, int shrink_log2) {
    RBS_PDL();
    int j = (int)RBS_TID;
    if (j >= n.NA) return;
    rbs_step_decide_item(n, dt, h_min, max_halvings, grow_iters, n_steps, shrink_log2, RBS_MEMBER(n, j));
}

// apply the decision to the states and the exact-forcing bookkeeping
__device__ __forceinline__ void rbs_step_apply_item(const Net& n, int predictor, int ent, int m) {
    int action = n.maction[m];
    if (action == ACTION_NONE) return;
    bool next = n.mphase[m] == PHASE_NEWTON;
    double h_next = n.mh[m], hl = n.mhlast[m];
    if (ent < n.n_state) {
        long long o = (long long)ent * n.B + m;
        double zl;
        if (action == ACTION_ACCEPT) {
            double un = n.u[o];
            zl = un - n.uprev[o];
            n.zlast[o] = zl;
            n.uprev[o] = un;
            if (n.fw_on && n.w_flow && n.mroll[m]) {
                // mean flow over the finished outer step = increment of the cumulative state
                // (save_flow, core/src/callback.jl:390-456)
                n.w_flow[(long long)n.n_state * n.B * (n.mstep[m] - 1) + o] = (un - n.ustep[o]) / n.fw_dt;
                n.ustep[o] = un;
            }
        } else {
            n.u[o] = n.uprev[o];
            zl = n.zlast[o];
        }
        if (next) n.z[o] = (predictor && hl > 0.0) ? zl * (h_next / hl) : 0.0;
        return;
    }
    const int ws = n.mstep[m], roll = n.mroll[m];
    const int b = ent - n.n_state;
    // the forcing of a finished window's last step keeps applying until the caller changes it
    rbs_exact_update(n, b, m, action == ACTION_ACCEPT, hl, next, h_next, ws - roll, next ? ws : ws - roll);
    if (n.fw_on && n.w_storage && roll)  // storages at the end of outer step ws - 1
        n.w_storage[n.fw_sb * (ws - 1) + (long long)b * n.B + m] = n.storage[(long long)b * n.B + m];
}

__global__ void k_step_apply(Net n, int predictor) {
    RBS_PDL();
    long long tid = RBS_TID;
    int ent = (int)(tid / n.NA);
    if (ent >= n.n_state + n.n_basin) return;
    rbs_step_apply_item(n, predictor, ent, RBS_MEMBER(n, (int)(tid - (long long)ent * n.NA)));
}

// ---- DiscreteControl (apply_discrete_control!, core/src/callback.jl:608-661) ----------------------
// One thread per (DiscreteControl node, member), after an accepted (sub-)step of that member:
// compound variables -> hysteresis conditions -> truth state -> control state.  Flows are the
// mean flows of the accepted step (u - uprev)/h, which is what get_du holds after an implicit
// Euler step; with `initial` (t = 0) the RHS du of the initial state is used and the control
// state is set even without a truth-state change.
__device__ __forceinline__ void rbs_discrete_control_item(const Net& n, int initial, int i, int m) {
    if (!initial && n.maction[m] != ACTION_ACCEPT) return;
    double inv_h = (!initial && n.mhlast[m] > 0.0) ? 1.0 / n.mhlast[m] : 0.0;
    int c0 = n.dc_cond_ptr[i], c1 = n.dc_cond_ptr[i + 1];
    int bits = 0;
    bool change = false;
    for (int c = c0; c < c1; c++) {
        double value = 0.0;
        for (int j = n.dc_csv_ptr[c]; j < n.dc_csv_ptr[c + 1]; j++) {
            int kind = n.dc_sv_kind[j], idx = n.dc_sv_idx[j];
            double x;
            if (kind == 0) x = n.level[(long long)idx * n.B + m];
            else if (kind == 1) x = n.storage[(long long)idx * n.B + m];
            else if (kind == 2) x = initial ? n.du[(long long)idx * n.B + m]
                                            : n.zlast[(long long)idx * n.B + m] * inv_h;
            else x = n.dc_sv_const[j];
            value += n.dc_sv_weight[j] * x;
        }
        long long o = (long long)c * n.B + m;
        int told = n.dc_truth[o];
        int tnew = told ? (value >= n.dc_thr_lo[c]) : (value > n.dc_thr_hi[c]);
        if (tnew != told) { change = true; n.dc_truth[o] = tnew; }
        bits |= tnew << (c - c0);
    }
    if (initial || change) {
        int snew = n.dc_logic[n.dc_logic_ptr[i] + bits];
        long long so = (long long)i * n.B + m;
        if (snew < 0) atomicOr(&n.status[m], 4);  // no control state specified for this truth state
        else if (snew != n.dc_state[so]) {
            n.dc_state[so] = snew;
            const int r = atomicAdd(&n.cnt_dc[m], 1);
            if (n.dc_rec_t && r < n.dc_rec_cap) {
                const long long ro = (long long)r * n.B + m;
                const double t_rel = initial ? 0.0
                    : (n.mroll[m] ? n.mstep[m] * n.win_dt : n.mstep[m] * n.win_dt + n.melapsed[m]);
                n.dc_rec_t[ro] = n.win_t0 + t_rel;
                n.dc_rec_ns[ro] = i | (snew << 16);
                n.dc_rec_bits[ro] = bits;
            }
        }
    }
}

__global__ void k_discrete_control(Net n, int initial) {
    RBS_PDL();
    long long tid = RBS_TID;
    int i = (int)(tid / n.NA);
    if (i >= n.n_dc) return;
    rbs_discrete_control_item(n, initial, i, RBS_MEMBER(n, (int)(tid - (long long)i * n.NA)));
}

// Host-driven steppers (the seam path: an integrator above the C ABI owns the step control): an
// accepted step of length dt moves the exactly integrated forcing forward
// (update_cumulative_flows!, core/src/callback.jl:131-170); same arithmetic as rbs_exact_update.
__global__ void k_accumulate_forcing(Net n, double dt) {
    RBS_PDL();
    long long tid = RBS_TID;
    const long long nb = (long long)n.n_basin * n.B;
    if (tid < nb) {
        const int b = (int)(tid / n.B);
        const double P = n.fixed_area[b] * n.precipitation[tid];
        n.cum_precipitation[tid] += P * dt;
        n.cum_surface_runoff[tid] += dt * n.surface_runoff[tid];
        n.cum_drainage[tid] += dt * n.drainage[tid];
        return;
    }
    tid -= nb;
    if (tid < (long long)n.n_fb * n.B) n.fb_cum[tid] += n.fb_rate[tid] * dt;
}

__global__ void k_fill(double* p, long long nelem, double v) {
    RBS_PDL();
    long long tid = RBS_TID;
    if (tid < nelem) p[tid] = v;
}
__global__ void k_broadcast(double* __restrict__ dst, const double* __restrict__ src, int n_entity, int B) {
    RBS_PDL();
    long long tid = RBS_TID;
    if (tid < (long long)n_entity * B) dst[tid] = src[tid / B];
}
