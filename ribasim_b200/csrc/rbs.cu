// rbs.cu — host side of librbs.so: the C ABI declared in include/ribasim_b200.h.
//
// Everything here is plumbing around the kernels in rbs_kernels.cuh: named device arrays,
// the `Net` kernel argument, launch sequencing of one RHS / Jacobian / linear solve / Newton
// step.  There is no CPU compute path: without a usable CUDA device rbs_create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/ribasim_b200.h"
#include "rbs_kernels.cuh"
#include "rbs_tile.cuh"
#include "rbs_fused.cuh"
#include "rbs_lane.cuh"

#define RBS_DC_REC_CAP 64  // DiscreteControl changes recorded per member (default; builder key dc_rec_cap)

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(e_));                   \
    } while (0)

struct DevArr {
    void* p = nullptr;
    size_t bytes = 0;
    bool is_int = false;
};

}  // namespace

struct rbs_builder {
    std::map<std::string, std::vector<int32_t>> i;
    std::map<std::string, std::vector<double>> d;
};

struct rbs_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int B = 0;
    Net net{};
    std::map<std::string, std::vector<int32_t>> hi;  // host copies of the integer descriptor
    std::map<std::string, DevArr> shared;            // descriptor + shared parameters
    std::map<std::string, DevArr> member;            // [n_entity][B] arrays
    std::map<std::string, int64_t> member_entities;
    double* staging = nullptr;  // device scratch for host<->device seam calls
    size_t staging_bytes = 0;
    int* h_active = nullptr;  // pinned: {n_active, n_limit}
    int *d_alist = nullptr, *d_llist = nullptr, *d_counts = nullptr;
    cudaStream_t cur = nullptr;  // stream the launch helpers use (the handle's, or a group's)
    // asynchronous host<->device staging: uploads land in shadow buffers on the copy stream and
    // become live at the start of the next step; downloads read a device-side snapshot
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_main = nullptr;
    std::map<std::string, void*> shadow_in, shadow_out;
    std::vector<std::string> staged;
    // forced windows (rbs_stage_forcing_window / rbs_advance_window): per forcing array a live and
    // a shadow buffer [steps][entity][B]; two result buffers so that the download of one window
    // overlaps with the computation of the next
    double *fw_live[6] = {}, *fw_shadow[6] = {};
    size_t fw_bytes[6] = {};
    int fw_staged_steps[6] = {}, fw_live_steps[6] = {};  // in slices
    int fw_every = 1, fw_phase = 0;                      // outer steps per slice, offset of step 0
    double* fw_res[2] = {nullptr, nullptr};
    size_t fw_res_bytes = 0;
    double* fw_flow[2] = {nullptr, nullptr};
    size_t fw_flow_bytes = 0;
    // progressive download of a running window's results: slices of outer steps that every member
    // has completed go out while the stragglers are still computing
    struct WinOut {
        bool on = false;
        double *host_s = nullptr, *dev_s = nullptr, *host_f = nullptr, *dev_f = nullptr;
        size_t slice_s = 0, slice_f = 0;
        int streamed = 0;
    } wout;
    int fw_res_cur = 0;
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_win = nullptr, ev_d2h[2] = {nullptr, nullptr};
    // member groups: independent state machines on their own streams, so the latency-bound
    // kernels of one group overlap with the kernels of the others
    struct Group {
        cudaStream_t stream = nullptr;
        cudaEvent_t ev = nullptr;
        int lo = 0, hi = 0;
        int *d_alist = nullptr, *d_llist = nullptr, *d_counts = nullptr;
        int* h_counts = nullptr;  // pinned {n_active, n_limit, outer steps completed by all members}
        int n_active = 0, n_limit = 0, min_step = 0;
        bool pending = false, done = false;
        int* d_tilecnt = nullptr;  // per 64-member tile: row blocks of k_newton_update done; [n_tiles]: k_limit_post_decide
        int n_tiles = 0;
    };
    std::vector<Group> groups;
    double tprev = 0.0;
    double abstol = 1e-5, reltol = 1e-5;
    int max_iter = 10;
    // stepper options (rbs_set_stepper)
    int full_newton = 1, predictor = 1, max_halvings = 20, grow_iters = 4;
    int early_exit = 1, shrink_log2 = 2;
    int64_t max_passes = 100000;
    int64_t launches = 0;
    int64_t n_steps = 0, n_passes = 0, n_failed = 0;
    // optional per-launch timing with CUDA events on the handle's stream (rbs_profile)
    bool profiling = false;
    struct ProfRec { const char* name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    // schedules (host)
    std::vector<int> lu_level_ptr, fwd_level_ptr, bwd_level_ptr;
    int *d_lu_level_ptr = nullptr, *d_fwd_level_ptr = nullptr, *d_bwd_level_ptr = nullptr;
    bool tile_mode = true;
    int tile = 8;
    int smem_tile = 0, smem_threads = 1024;  // shared-memory factor+solve kernel (0 = does not fit)
    LinSched sched{};
    LduSched dsched{};
    void* d_blob2 = nullptr;
    bool lin_smem16 = false;
    bool lin_smem = false;   // factors in shared memory (tile 8) instead of L2 (tile 8/16/32)
    size_t lin_smem_bytes = 0;
    bool sched16 = false;
    // entry schedule (factorisation with fused forward substitution), see symbolic.build_entries
    EntSched esched{};
    EntSched esched_c{};         // compact slots (full-Newton path): more tiles per SM
    void* d_entc_blob = nullptr;
    size_t entc_smem_bytes = 0;
    bool lin_entc = false;
    double limit_frac = 0.0;     // see launch_pass
    int lin_lane = 0;            // member-per-lane solve on global slots (rbs_lane.cuh): warps per block, 0 = off
    int lane_min = 2048;         // launches of at least this many members take it (measured crossover with the shared-memory kernel)
    LaneTape lane_tape{};
    void* d_lane_blob = nullptr;
    size_t lane_smem = 0;
    // M members per thread on the compact entry schedule (rbs_fused.cuh); 0 = off
    int lin_m = 0;
    int lin_m_sched = 0;         // schedule words in shared memory behind the slots
    size_t lin_m_smem = 0;
    // round-2 pass: convergence test + finalize half as one cluster kernel at the END of the pass
    // in which a member's iteration ended (rbs_fused.cuh); 0 = the round-1 launch sequence
    bool pdl = true;             // programmatic dependent launch of the step-path kernels
    int r2_pass = 0;
    int fin_cluster = 4;
    void* d_ent_blob = nullptr;
    bool lin_ent = false;
    int ent_tile = 2;
    size_t ent_smem_bytes = 0;
    void* d_blob = nullptr;
    size_t smem_bytes = 0;
    int *d_lsrc_ptr = nullptr, *d_lsrc_code = nullptr;
    std::vector<int> list_phase[3];
    int dc_rec_cap = RBS_DC_REC_CAP;  // DiscreteControl changes recorded per member
    bool pid_fused = false;  // k_pid also formulates its controlled pump / outlet
    bool pid_kd_nonzero = false;  // a PidControl derivative gain is in force (its D-term reads flows)
    bool pid_cs_kd_nonzero = false;  // ... in a control-state parameter row
    bool fuse_small = true;       // fewer, fused launches per pass (RBS_FUSE=0: the plain sequence)
    int iters_per_pass = 1;       // Newton iterations launched per pass (measured: 1 is best)
    bool host_threads = false;    // one host thread per member group in the pass loop (measured: no gain)
    // cooperative window kernel (rbs_coop.cuh): used when the ensemble is small
    int coop_members = 0;    // use it for ensembles of at most this many members (0 = never)
    int coop_blocks = 0;     // co-resident grid size (0 = unavailable)
    int tile_window = 0;     // members per block of the tile window kernel (0 = stream path)
    size_t tile_smem = 0;
    int* d_coop_counters = nullptr;
    int* h_coop_counters = nullptr;  // pinned
    int* d_list_phase[3] = {nullptr, nullptr, nullptr};
    int n_part = 1, rows_per_block = 64;
    bool factor_valid = false;
};

namespace {

// Launch state of the calling thread.  The pass loop of a window runs one host thread per member
// group; each has its own copy of the kernel argument `Net` (launch domain, work list) and its own
// stream, so that the helpers below can be shared.  Without a context the handle's own apply.
struct LaunchCtx {
    Net net;
    cudaStream_t cur = nullptr;
    int64_t launches = 0, n_passes = 0;
};
thread_local LaunchCtx* tl_ctx = nullptr;
inline Net& net_of(rbs_handle* h) { return tl_ctx ? tl_ctx->net : h->net; }
inline cudaStream_t& cur_of(rbs_handle* h) { return tl_ctx ? tl_ctx->cur : h->cur; }
inline int64_t& launches_of(rbs_handle* h) { return tl_ctx ? tl_ctx->launches : h->launches; }
inline int64_t& passes_of(rbs_handle* h) { return tl_ctx ? tl_ctx->n_passes : h->n_passes; }

inline int grid_for(long long total, int block = 256) { return (int)((total + block - 1) / block); }

template <class T> int upload(rbs_handle* h, const std::string& key, const std::vector<T>& v, bool is_int) {
    DevArr& a = h->shared[key];
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    if (a.bytes != bytes) {
        if (a.p) cudaFree(a.p);
        CUDA_TRY(cudaMalloc(&a.p, bytes));
        a.bytes = bytes;
        a.is_int = is_int;
    }
    if (!v.empty())
        CUDA_TRY(cudaMemcpyAsync(a.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int alloc_member(rbs_handle* h, const std::string& key, int64_t n_entity, bool is_int = false) {
    DevArr a;
    size_t elem = is_int ? sizeof(int) : sizeof(double);
    a.bytes = std::max<size_t>((size_t)n_entity * h->B, 1) * elem;
    a.is_int = is_int;
    CUDA_TRY(cudaMalloc(&a.p, a.bytes));
    CUDA_TRY(cudaMemsetAsync(a.p, 0, a.bytes, h->stream));
    h->member[key] = a;
    h->member_entities[key] = n_entity;
    return 0;
}

template <class T> T* sp(rbs_handle* h, const char* key) {
    auto it = h->shared.find(key);
    return it == h->shared.end() ? nullptr : (T*)it->second.p;
}
double* mp(rbs_handle* h, const char* key) {
    auto it = h->member.find(key);
    return it == h->member.end() ? nullptr : (double*)it->second.p;
}

int geti(const rbs_builder* b, const char* key, int dflt = 0) {
    auto it = b->i.find(key);
    return it == b->i.end() || it->second.empty() ? dflt : it->second[0];
}

void bind_net(rbs_handle* h) {
    Net& n = h->net;
#define S_D(name) n.name = sp<double>(h, #name)
#define S_I(name) n.name = sp<int>(h, #name)
#define M_D(name) n.name = mp(h, #name)
    S_D(storage0); S_D(low_storage_threshold); S_D(fixed_area);
    S_I(prof_ptr); S_D(prof_level); S_D(prof_dSdh); S_D(prof_dSdh_slope); S_D(prof_cumS);
    S_D(prof_area); S_D(prof_area_slope);
    S_I(inc_ptr); S_I(inc_state); S_D(inc_sign); S_I(bfb_ptr); S_I(bfb_idx); S_I(basin_obs);
    S_I(conn_type); S_I(conn_idx); S_I(src_kind); S_I(src_idx); S_I(dst_kind); S_I(dst_idx);
    S_D(lr_resistance); S_D(lr_max_flow_rate);
    S_D(mn_length); S_D(mn_manning_n); S_D(mn_profile_width); S_D(mn_profile_slope);
    S_D(mn_upstream_bottom); S_D(mn_downstream_bottom);
    S_I(tab_ptr); S_D(tab_t); S_D(tab_u); S_D(tab_du); S_D(tab_c1); S_D(tab_c2);
    S_D(tab_slope_left); S_D(tab_slope_right);
    S_I(pump_control_type); S_I(outlet_control_type);
    S_I(ud_link_ptr); S_I(ud_of_link); S_D(ud_min_level); S_D(ud_link_alloc);
    S_D(lb_level); S_I(trc_table_idx); S_D(trc_max_downstream_level);
    S_D(pump_flow_rate); S_D(pump_min_flow_rate); S_D(pump_max_flow_rate);
    S_D(pump_min_upstream_level); S_D(pump_max_downstream_level); S_D(pump_flow_demand);
    S_D(outlet_flow_rate); S_D(outlet_min_flow_rate); S_D(outlet_max_flow_rate);
    S_D(outlet_min_upstream_level); S_D(outlet_max_downstream_level); S_D(outlet_flow_demand);
    S_D(ud_total_demand); S_D(ud_return_factor);
    S_D(pid_target); S_D(pid_proportional); S_D(pid_integral); S_D(pid_derivative);
    S_D(cc_sv_const);
    S_I(pid_listen_basin); S_I(pid_target_is_outlet); S_I(pid_target_idx); S_I(pid_target_state);
    S_I(pid_jpos_listen); S_I(pid_jpos_integral); S_I(pid_int_jpos);
    S_I(pid_term_ptr); S_I(pid_term_state); S_D(pid_term_sign);
    S_I(pid_jterm_ptr); S_I(pid_jterm_src); S_I(pid_jterm_dst); S_D(pid_jterm_sign);
    S_I(cc_target_is_outlet); S_I(cc_target_idx); S_I(cc_target_state); S_I(cc_sv_ptr);
    S_I(cc_sv_kind); S_I(cc_sv_idx); S_D(cc_sv_weight);
    S_I(cc_term_ptr); S_I(cc_term_kind); S_I(cc_term_sv); S_I(cc_term_src); S_I(cc_term_dst);
    S_I(dc_cond_ptr); S_I(dc_csv_ptr); S_I(dc_sv_kind); S_I(dc_sv_idx); S_I(dc_logic_ptr); S_I(dc_logic);
    S_D(dc_sv_weight); S_D(dc_thr_hi); S_D(dc_thr_lo); S_D(dc_sv_const);
    S_I(pump_dc); S_I(pump_cs_ptr); S_I(outlet_dc); S_I(outlet_cs_ptr); S_I(lr_dc); S_I(lr_cs_ptr);
    S_I(trc_dc); S_I(trc_cs_ptr); S_I(trc_cs_table);
    S_D(pump_cs_flow_rate); S_D(pump_cs_min_flow_rate); S_D(pump_cs_max_flow_rate);
    S_D(pump_cs_min_upstream_level); S_D(pump_cs_max_downstream_level);
    S_D(outlet_cs_flow_rate); S_D(outlet_cs_min_flow_rate); S_D(outlet_cs_max_flow_rate);
    S_D(outlet_cs_min_upstream_level); S_D(outlet_cs_max_downstream_level);
    S_D(lr_cs_resistance); S_D(lr_cs_max_flow_rate);
    S_I(mn_dc); S_I(mn_cs_ptr); S_I(pid_dc); S_I(pid_cs_ptr);
    S_D(mn_cs_length); S_D(mn_cs_manning_n); S_D(mn_cs_profile_width); S_D(mn_cs_profile_slope);
    S_D(pid_cs_target); S_D(pid_cs_proportional); S_D(pid_cs_integral); S_D(pid_cs_derivative);
    S_I(jrow_ptr); S_I(jcol); S_I(jpos_src); S_I(jpos_dst); S_I(ud_out_jpos);
    S_I(wsrc_ptr); S_I(wsrc_pos); S_I(wdiag_flag); S_D(wsrc_sign);
    S_I(lu_wsrc); S_I(lu_pivot); S_I(lu_prod_ptr); S_I(lu_prod_l); S_I(lu_prod_u); S_I(lu_diag);
    S_I(perm); S_I(fwd_row); S_I(fwd_ptr); S_I(fwd_l); S_I(fwd_k); S_I(bwd_row); S_I(bwd_ptr);
    S_I(bwd_u); S_I(bwd_k);
    M_D(u); M_D(uprev); M_D(z); M_D(du); M_D(ur); M_D(br); M_D(xs);
    M_D(storage); M_D(level); M_D(factor); M_D(area); M_D(dlevel); M_D(dfactor); M_D(darea);
    M_D(precipitation); M_D(surface_runoff); M_D(potential_evaporation); M_D(drainage);
    M_D(infiltration); M_D(cum_precipitation); M_D(cum_surface_runoff); M_D(cum_drainage);
    M_D(exact); M_D(fb_rate); M_D(fb_cum); M_D(frp); M_D(fro); M_D(jv); M_D(wv); M_D(lu);
    M_D(norm_part); M_D(ndz); M_D(ndz_prev); M_D(eta); M_D(storage_prev); M_D(level_prev);
    M_D(mh); M_D(mhnom); M_D(melapsed); M_D(mhlast); M_D(zlast);
    n.pm_mn_manning_n = mp(h, "pm_mn_manning_n"); n.pm_lr_resistance = mp(h, "pm_lr_resistance");
    n.pm_pump_flow_rate = mp(h, "pm_pump_flow_rate"); n.pm_outlet_flow_rate = mp(h, "pm_outlet_flow_rate");
    n.pm_lb_level = mp(h, "pm_lb_level");
#define M_I(name) n.name = (int*)h->member[#name].p
    M_I(nstat); M_I(status); M_I(active_count);
    M_I(mphase); M_I(mjac); M_I(miter); M_I(mconv); M_I(mneg); M_I(mneg2); M_I(mclamp); M_I(mlast); M_I(maction);
    M_I(cnt_iters); M_I(cnt_sub); M_I(cnt_rej); M_I(mstep); M_I(mroll);
    M_I(dc_truth); M_I(dc_state); M_I(cnt_dc);
    n.dc_rec_cap = n.n_dc > 0 ? h->dc_rec_cap : 0;
    n.dc_rec_t = n.n_dc > 0 ? mp(h, "dc_rec_t") : nullptr;
    n.dc_rec_ns = n.n_dc > 0 ? (int*)h->member["dc_rec_ns"].p : nullptr;
    n.dc_rec_bits = n.n_dc > 0 ? (int*)h->member["dc_rec_bits"].p : nullptr;
#undef M_I
#undef S_D
#undef S_I
#undef M_D
}

inline void prof_begin(rbs_handle* h, const char* name) {
    if (!h->profiling) return;
    rbs_handle::ProfRec r{name, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, h->cur);
    h->prof.push_back(r);
}
inline void prof_end(rbs_handle* h) {
    if (h->profiling) cudaEventRecord(h->prof.back().b, h->cur);
}

// launch with the programmatic-dependent-launch attribute (see RBS_PDL in rbs_kernels.cuh)
template <class... KArgs, class... Args>
inline void launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define LAUNCH_CFG_AS(h, label, kernel, grid, block, ...)                                      \
    do {                                                                                       \
        prof_begin((h), label);                                                                \
        launch_k((h)->pdl && !(h)->profiling, kernel, dim3(grid), dim3(block), 0, cur_of(h), __VA_ARGS__); \
        prof_end(h);                                                                           \
        launches_of(h)++;                                                                      \
    } while (0)
#define LAUNCH_CFG(h, kernel, grid, block, ...) LAUNCH_CFG_AS(h, #kernel, kernel, grid, block, __VA_ARGS__)

// grid over (entities x launch domain), profile label = kernel name or an explicit one
#define LAUNCH_AS(h, label, kernel, total, ...)                                                \
    do {                                                                                       \
        long long tot_ = (total);                                                              \
        if (tot_ > 0) LAUNCH_CFG_AS(h, label, kernel, grid_for(tot_), 256, __VA_ARGS__);       \
    } while (0)
#define LAUNCH(h, kernel, total, ...) LAUNCH_AS(h, #kernel, kernel, total, __VA_ARGS__)

// RHS phases on the caches prepared by k_exact / k_step_apply: ur -> du (and J values when JAC).
// SRC 0 reads u_reduced from `ur`; SRC 1 fuses reduce_state!(uprev + z) into the basin kernel.
template <bool JAC, int SRC> void launch_rhs(rbs_handle* h, const double* ur, Pred pred = Pred{nullptr, 0, nullptr}) {
    Net& n = net_of(h);
    long long B = n.NA;
    LAUNCH_AS(h, JAC ? "k_basin<jac>" : "k_basin<rhs>", (k_basin<JAC, SRC>), (long long)n.n_reduced * B, n, ur, true, pred, 0);
    // the PidControl items ride in the links launch when nothing reads a flow in between
    n.pid_in_links = (SRC == 1 && h->pid_fused && !h->pid_kd_nonzero && n.n_cc == 0 && h->fuse_small) ? 1 : 0;
    LAUNCH_AS(h, JAC ? "k_links<jac>" : "k_links<rhs>", k_links<JAC>, (long long)(n.n_conn + (n.pid_in_links ? n.n_pid : 0)) * B, n, 0,
              nullptr, 0, pred);
    if (n.n_cc > 0) {
        LAUNCH_AS(h, JAC ? "k_continuous_control<jac>" : "k_continuous_control<rhs>", k_continuous_control<JAC>, (long long)n.n_cc * B, n, pred);
        LAUNCH_AS(h, JAC ? "k_links_cc<jac>" : "k_links_cc<rhs>", k_links<JAC>, (long long)h->list_phase[1].size() * B, n, 1, h->d_list_phase[1],
                  (int)h->list_phase[1].size(), pred);
    }
    if (n.n_pid > 0 && !n.pid_in_links) {
        LAUNCH_AS(h, JAC ? "k_pid<jac>" : "k_pid<rhs>", k_pid<JAC>, (long long)n.n_pid * B, n, n.ur, pred, h->pid_fused ? 1 : 0);
        if (!h->pid_fused)
            LAUNCH_AS(h, JAC ? "k_links_pid<jac>" : "k_links_pid<rhs>", k_links<JAC>, (long long)h->list_phase[2].size() * B, n, 2, h->d_list_phase[2],
                      (int)h->list_phase[2].size(), pred);
    }
}
// W = A J - I/gamma and its LU.  FROM_J: assembled on the fly (step path, per-member gamma = h).
template <bool FROM_J> void launch_factor(rbs_handle* h, double inv_gamma, Pred pred = Pred{nullptr, 0, nullptr}) {
    Net& n = net_of(h);
    long long B = n.NA;
    if (!FROM_J) LAUNCH(h, k_assemble_w, (long long)n.nnz_w * B, n, inv_gamma);
    int n_levels = (int)h->lu_level_ptr.size() - 1;
    if (h->tile_mode) {
        int blocks = (n.NA + h->tile - 1) / h->tile;
        LAUNCH_CFG_AS(h, "k_lu_tile", k_lu_tile<FROM_J>, blocks, 256, n, h->d_lu_level_ptr, n_levels, h->tile, inv_gamma, pred);
    } else {
        for (int lv = 0; lv < n_levels; lv++) {
            int e0 = h->lu_level_ptr[lv], e1 = h->lu_level_ptr[lv + 1];
            LAUNCH_AS(h, "k_lu_level", k_lu_level<FROM_J>, (long long)(e1 - e0) * B, n, e0, e1, inv_gamma, pred);
        }
    }
    h->factor_valid = true;
}

template <bool FROM_DU> void launch_solve(rbs_handle* h, const double* b, double* x, Pred pred = Pred{nullptr, 0, nullptr}) {
    Net& n = net_of(h);
    long long B = n.NA;
    int n_fwd = (int)h->fwd_level_ptr.size() - 1, n_bwd = (int)h->bwd_level_ptr.size() - 1;
    if (h->tile_mode) {
        int blocks = (n.NA + h->tile - 1) / h->tile;
        LAUNCH_CFG_AS(h, "k_solve_tile", k_solve_tile<FROM_DU>, blocks, 256, n, b, x, h->d_fwd_level_ptr, n_fwd,
                   h->d_bwd_level_ptr, n_bwd, h->tile, pred);
    } else {
        LAUNCH_AS(h, "k_solve_load", k_solve_load<FROM_DU>, (long long)n.n_reduced * B, n, b, pred);
        for (int lv = 1; lv < n_fwd; lv++) {
            int p0 = h->fwd_level_ptr[lv], p1 = h->fwd_level_ptr[lv + 1];
            LAUNCH(h, k_solve_level, (long long)(p1 - p0) * B, n, p0, p1, 0, pred);
        }
        for (int lv = 0; lv < n_bwd; lv++) {
            int p0 = h->bwd_level_ptr[lv], p1 = h->bwd_level_ptr[lv + 1];
            LAUNCH(h, k_solve_level, (long long)(p1 - p0) * B, n, p0, p1, 1, pred);
        }
        LAUNCH(h, k_solve_store, (long long)n.n_reduced * B, n, x, pred);
    }
}

int ensure_staging(rbs_handle* h, size_t bytes) {
    if (h->staging_bytes >= bytes) return 0;
    if (h->staging) cudaFree(h->staging);
    h->staging = nullptr;
    h->staging_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->staging, bytes));
    h->staging_bytes = bytes;
    return 0;
}

void free_groups(rbs_handle* h) {
    for (auto& G : h->groups) {
        if (G.d_alist) cudaFree(G.d_alist);
        if (G.d_llist) cudaFree(G.d_llist);
        if (G.d_counts) cudaFree(G.d_counts);
        if (G.h_counts) cudaFreeHost(G.h_counts);
        if (G.d_tilecnt) cudaFree(G.d_tilecnt);
        if (G.ev) cudaEventDestroy(G.ev);
        if (G.stream && G.stream != h->stream) cudaStreamDestroy(G.stream);
    }
    h->groups.clear();
}

// Partition the members into `ng` groups (0 = auto) with their own streams and work lists.
int make_groups(rbs_handle* h, int ng) {
    free_groups(h);
    int n_members = h->B;
    if (ng <= 0) ng = n_members >= 2048 ? 4 : (n_members >= 512 ? 2 : 1);
    ng = std::min(ng, std::max(1, n_members / 32));
    h->groups.resize(ng);
    // group boundaries on multiples of 32 members keep every group's accesses aligned
    int per = ((n_members + ng - 1) / ng + 31) / 32 * 32;
    for (int g = 0; g < ng; g++) {
        auto& G = h->groups[g];
        G.lo = std::min(g * per, n_members);
        G.hi = std::min(G.lo + per, n_members);
        if (g == 0) G.stream = h->stream;
        else CUDA_TRY(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&G.ev, cudaEventDisableTiming));
        size_t nb_ = sizeof(int) * std::max(G.hi - G.lo, 1);
        CUDA_TRY(cudaMalloc(&G.d_alist, nb_));
        CUDA_TRY(cudaMalloc(&G.d_llist, nb_));
        CUDA_TRY(cudaMalloc(&G.d_counts, 4 * sizeof(int)));
        CUDA_TRY(cudaMallocHost(&G.h_counts, 4 * sizeof(int)));
        const size_t nt = (size_t)(std::max(G.hi - G.lo, 1) + 63) / 64;
        G.n_tiles = (int)nt;
        CUDA_TRY(cudaMalloc(&G.d_tilecnt, (nt + 1) * sizeof(int)));
        CUDA_TRY(cudaMemset(G.d_tilecnt, 0, (nt + 1) * sizeof(int)));
    }
    return 0;
}

int check_async(rbs_handle* h, const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string(where) + ": " + cudaGetErrorString(e));
    return 0;
}

}  // namespace

extern "C" {

const char* rbs_version(void) { return "ribasim_b200 0.1.0 (sm_100a)"; }

int rbs_last_error(char* buf, int32_t len) {
    if (buf && len > 0) {
        strncpy(buf, g_err.c_str(), (size_t)len - 1);
        buf[len - 1] = 0;
    }
    return (int)g_err.size();
}

rbs_builder* rbs_builder_create(void) { return new rbs_builder(); }
void rbs_builder_destroy(rbs_builder* b) { delete b; }
int rbs_builder_set_i32(rbs_builder* b, const char* key, const int32_t* data, int64_t n) {
    if (!b || !key || (n > 0 && !data)) return fail("rbs_builder_set_i32: null argument");
    b->i[key].assign(data, data + n);
    return 0;
}
int rbs_builder_set_f64(rbs_builder* b, const char* key, const double* data, int64_t n) {
    if (!b || !key || (n > 0 && !data)) return fail("rbs_builder_set_f64: null argument");
    b->d[key].assign(data, data + n);
    return 0;
}

// Shared-memory carve-out of the entry-schedule solve.  What is not carved out is L1, and the
// schedule words (47 KB for the 500-basin network) are read through L1 by every tile: measured on
// the bench workload (profiles/compact_slots.sh), three resident tiles with the rest as L1 beat
// four to six tiles with a small L1, alone and next to the other member groups' kernels.  The
// preference is therefore three tiles' worth, for the slot layout the current stepper mode uses.
static void apply_carveout(rbs_handle* h) {
    if (!h->lin_ent) return;
    const bool cmp = h->lin_entc && h->full_newton;
    const size_t tile = (cmp ? h->entc_smem_bytes : h->ent_smem_bytes) + 1040;
    // the carve-out sizes of sm_100 (KB); the preference is given in percent of the largest
    int kb = 228;
    for (int cand : {8, 16, 32, 64, 100, 132, 164, 196, 228})
        if ((size_t)cand * 1024 >= 3 * tile) { kb = cand; break; }
    int pct = (kb * 100 + 227) / 228;
    if (getenv("RBS_ENT_CARVEOUT")) pct = atoi(getenv("RBS_ENT_CARVEOUT"));
    if (pct < 0) return;  // RBS_ENT_CARVEOUT=-1: the driver's default
    if (pct > 100) pct = 100;
    if (h->ent_tile == 2) cudaFuncSetAttribute(k_newton_linear_ent<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    else if (h->ent_tile == 4) cudaFuncSetAttribute(k_newton_linear_ent<4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    else cudaFuncSetAttribute(k_newton_linear_ent<8>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaGetLastError();
}

rbs_handle* rbs_create(const rbs_builder* b, int32_t n_members, int32_t device) {
    if (!b) { fail("rbs_create: null builder"); return nullptr; }
    if (n_members < 1) { fail("rbs_create: n_members must be >= 1"); return nullptr; }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        fail(std::string("rbs_create: no usable CUDA device (") + cudaGetErrorString(e) +
             "); librbs has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= n_dev) { fail("rbs_create: invalid device index"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail("rbs_create: cudaSetDevice failed"); return nullptr; }
    rbs_handle* h = new rbs_handle();
    h->device = device;
    h->B = n_members;
    auto bail = [&](const char* why) -> rbs_handle* {
        if (g_err.empty()) fail(why);
        rbs_destroy(h);
        return nullptr;
    };
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail("rbs_create: stream creation failed");
    h->cur = h->stream;
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess)
        return bail("rbs_create: copy stream creation failed");
    if (cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_h2d, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_win, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_d2h[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_d2h[1], cudaEventDisableTiming) != cudaSuccess)
        return bail("rbs_create: window stream creation failed");
    if (cudaMallocHost(&h->h_active, 2 * sizeof(int)) != cudaSuccess) return bail("rbs_create: pinned alloc failed");
    if (make_groups(h, geti(b, "n_groups", 0))) return bail("rbs_create: group setup failed");
    for (auto& kv : b->i) { h->hi[kv.first] = kv.second; if (upload(h, kv.first, kv.second, true)) return bail("upload"); }
    for (auto& kv : b->d) if (upload(h, kv.first, kv.second, false)) return bail("upload");
    Net& n = h->net;
    n.B = n_members;
    n.NA = n_members;
    n.alist = nullptr;
    const char* required[] = {"n_basin", "n_state", "n_conn", "n_reduced", "nnz_j", "nnz_w", "n_lu"};
    for (const char* k : required)
        if (b->i.find(k) == b->i.end()) { fail(std::string("rbs_create: missing descriptor key ") + k); return bail(""); }
#define GI(name) n.name = geti(b, #name)
    GI(n_basin); GI(n_pid); GI(n_state); GI(n_conn); GI(n_reduced); GI(n_trc); GI(n_pump);
    GI(n_outlet); GI(n_ud); GI(n_ud_link); GI(n_lr); GI(n_manning); GI(n_lb); GI(n_fb); GI(n_cc);
    GI(n_tab); GI(off_trc); GI(off_pump); GI(off_outlet); GI(off_ud_in); GI(off_ud_out); GI(off_lr);
    GI(off_manning); GI(off_evap); GI(off_infil); GI(off_integral); GI(nnz_j); GI(nnz_w); GI(n_lu);
    GI(cc_tab_offset); GI(n_dc); GI(n_dc_cond);
#undef GI
    {
        auto it = b->d.find("level_difference_threshold");
        n.ldt = it == b->d.end() || it->second.empty() ? 0.02 : it->second[0];
    }
    // shared time-dependent parameters: allocate with neutral defaults
    struct PD { const char* key; int count; double dflt; };
    const double inf = INFINITY, nan = NAN;
    PD pds[] = {
        {"lb_level", n.n_lb, 0}, {"trc_max_downstream_level", n.n_trc, inf},
        {"pump_flow_rate", n.n_pump, 0}, {"pump_min_flow_rate", n.n_pump, 0},
        {"pump_max_flow_rate", n.n_pump, inf}, {"pump_min_upstream_level", n.n_pump, -inf},
        {"pump_max_downstream_level", n.n_pump, inf}, {"pump_flow_demand", n.n_pump, nan},
        {"outlet_flow_rate", n.n_outlet, 0}, {"outlet_min_flow_rate", n.n_outlet, 0},
        {"outlet_max_flow_rate", n.n_outlet, inf}, {"outlet_min_upstream_level", n.n_outlet, -inf},
        {"outlet_max_downstream_level", n.n_outlet, inf}, {"outlet_flow_demand", n.n_outlet, nan},
        {"ud_total_demand", n.n_ud, 0}, {"ud_return_factor", n.n_ud, 0},
        {"ud_alloc_total", n.n_ud, 0},
        {"pid_target", n.n_pid, 0}, {"pid_proportional", n.n_pid, 0}, {"pid_integral", n.n_pid, 0},
        {"pid_derivative", n.n_pid, 0},
        {"dc_thr_hi", n.n_dc_cond, inf}, {"dc_thr_lo", n.n_dc_cond, inf},
    };
    for (auto& pd : pds)
        if (h->shared.find(pd.key) == h->shared.end())
            if (upload(h, pd.key, std::vector<double>((size_t)pd.count, pd.dflt), false)) return bail("param alloc");
    if (h->shared.find("trc_table_idx") == h->shared.end())
        if (upload(h, "trc_table_idx", std::vector<int32_t>((size_t)n.n_trc, 0), true)) return bail("param alloc");
    if (h->shared.find("ud_demand_from_timeseries") == h->shared.end())
        if (upload(h, "ud_demand_from_timeseries", std::vector<int32_t>((size_t)n.n_ud, 0), true)) return bail("param alloc");
    if (h->shared.find("dc_sv_const") == h->shared.end()) {
        size_t nsv = h->hi.count("dc_sv_kind") ? h->hi["dc_sv_kind"].size() : 0;
        if (upload(h, "dc_sv_const", std::vector<double>(nsv, 0.0), false)) return bail("param alloc");
    }
    if (h->shared.find("cc_sv_const") == h->shared.end()) {
        size_t nsv = h->hi.count("cc_sv_kind") ? h->hi["cc_sv_kind"].size() : 0;
        if (upload(h, "cc_sv_const", std::vector<double>(nsv, 0.0), false)) return bail("param alloc");
    }
    {
        // basins whose storage / area a controller reads while the RHS is evaluated
        std::vector<int32_t> obs((size_t)std::max(n.n_basin, 1), 0);
        auto mark = [&](const char* key) { auto it = h->hi.find(key); if (it != h->hi.end()) for (int32_t b_ : it->second) if (b_ >= 0 && b_ < n.n_basin) obs[b_] = 1; };
        mark("pid_listen_basin");
        auto ik = h->hi.find("cc_sv_kind"), ii = h->hi.find("cc_sv_idx");
        if (ik != h->hi.end() && ii != h->hi.end())
            for (size_t k = 0; k < ik->second.size() && k < ii->second.size(); k++)
                if (ik->second[k] <= 1 && ii->second[k] >= 0 && ii->second[k] < n.n_basin) obs[ii->second[k]] = 1;
        if (upload(h, "basin_obs", obs, true)) return bail("param alloc");
    }
    // per-member arrays
    // rows of the state vector per block of k_newton_update (a thread walks rows_per_block / 8 rows one
    // after the other): fixed for every ensemble size, because the partition fixes the summation order
    // of the convergence norm and member results must not depend on the batch
    h->limit_frac = getenv("RBS_LIMIT_FRAC") ? std::min(1.0, std::max(0.0, atof(getenv("RBS_LIMIT_FRAC")))) : 0.0;
    h->rows_per_block = getenv("RBS_ROWS_PER_BLOCK") ? std::max(8, atoi(getenv("RBS_ROWS_PER_BLOCK"))) : geti(b, "rows_per_block", 64);
    h->n_part = (n.n_state + h->rows_per_block - 1) / h->rows_per_block;
    struct MD { const char* key; int64_t n; };
    MD mds[] = {
        {"u", n.n_state}, {"uprev", n.n_state}, {"z", n.n_state}, {"du", n.n_state},
        {"ur", n.n_reduced}, {"br", n.n_reduced}, {"xs", n.n_reduced}, {"cvec", n.n_reduced},
        {"storage", n.n_basin}, {"level", n.n_basin}, {"factor", n.n_basin}, {"area", n.n_basin},
        {"dlevel", n.n_basin}, {"dfactor", n.n_basin}, {"darea", n.n_basin},
        {"precipitation", n.n_basin}, {"surface_runoff", n.n_basin},
        {"potential_evaporation", n.n_basin}, {"drainage", n.n_basin}, {"infiltration", n.n_basin},
        {"cum_precipitation", n.n_basin}, {"cum_surface_runoff", n.n_basin},
        {"cum_drainage", n.n_basin}, {"exact", n.n_basin}, {"storage_prev", n.n_basin},
        {"level_prev", n.n_basin}, {"fb_rate", n.n_fb}, {"fb_cum", n.n_fb}, {"frp", n.n_pump},
        {"fro", n.n_outlet}, {"jv", n.nnz_j}, {"wv", n.nnz_w},
        // factor slots; the rows behind n_lu hold the right-hand-side column of the member-per-lane solve
        {"lu", n.n_lu + n.n_reduced},
        {"norm_part", h->n_part}, {"ndz", 1}, {"ndz_prev", 1}, {"eta", 1},
        {"zlast", n.n_state}, {"ustep", n.n_state}, {"mh", 1}, {"mhnom", 1}, {"melapsed", 1}, {"mhlast", 1},
    };
    for (auto& md : mds) if (alloc_member(h, md.key, md.n)) return bail("member alloc");
    h->member_entities["lu"] = n.n_lu;
    for (const char* k : {"nstat", "status", "mphase", "mjac", "miter", "mconv", "mneg", "mneg2", "mclamp", "mlast",
                          "maction", "cnt_iters", "cnt_sub", "cnt_rej", "mstep", "mroll", "cnt_dc"})
        if (alloc_member(h, k, 1, true)) return bail("member alloc");
    if (alloc_member(h, "dc_truth", n.n_dc_cond, true) || alloc_member(h, "dc_state", n.n_dc, true))
        return bail("member alloc");
    h->dc_rec_cap = std::max(1, geti(b, "dc_rec_cap", RBS_DC_REC_CAP));
    if (n.n_dc > 0 && (alloc_member(h, "dc_rec_t", h->dc_rec_cap) || alloc_member(h, "dc_rec_ns", h->dc_rec_cap, true) ||
                       alloc_member(h, "dc_rec_bits", h->dc_rec_cap, true)))
        return bail("member alloc");
    // no control state set yet
    if (cudaMemsetAsync(h->member["dc_state"].p, 0xFF, h->member["dc_state"].bytes, h->stream) != cudaSuccess)
        return bail("dc_state init");
    {
        DevArr a; a.bytes = sizeof(int); a.is_int = true;
        if (cudaMalloc(&a.p, a.bytes) != cudaSuccess) return bail("active_count alloc");
        if (cudaMemsetAsync(a.p, 0, a.bytes, h->stream) != cudaSuccess) return bail("active_count init");
        h->member["active_count"] = a; h->member_entities["active_count"] = 0;
    }
    // eta_old starts at 1 (nlsolver.ηold)
    k_fill<<<grid_for(n.B), 256, 0, h->stream>>>(mp(h, "eta"), n.B, 1.0);
    // storage_prev = storage0 (core/src/read.jl:944-946)
    k_broadcast<<<grid_for((long long)n.n_basin * n.B), 256, 0, h->stream>>>(
        mp(h, "storage_prev"), sp<double>(h, "storage0"), n.n_basin, n.B);
    // schedules
    auto getv = [&](const char* k) { auto it = b->i.find(k); return it == b->i.end() ? std::vector<int32_t>{0} : it->second; };
    h->lu_level_ptr = getv("lu_level_ptr");
    h->fwd_level_ptr = getv("fwd_level_ptr");
    h->bwd_level_ptr = getv("bwd_level_ptr");
    h->d_lu_level_ptr = sp<int>(h, "lu_level_ptr");
    h->d_fwd_level_ptr = sp<int>(h, "fwd_level_ptr");
    h->d_bwd_level_ptr = sp<int>(h, "bwd_level_ptr");
    // controlled pump/outlet states per phase
    auto& pct = h->hi["pump_control_type"];
    auto& oct = h->hi["outlet_control_type"];
    for (int k = 0; k < n.n_pump; k++) if (pct[k] != CONTROL_NONE) h->list_phase[pct[k]].push_back(n.off_pump + k);
    for (int k = 0; k < n.n_outlet; k++) if (oct[k] != CONTROL_NONE) h->list_phase[oct[k]].push_back(n.off_outlet + k);
    {
        // The controlled link can be formulated by the controller's own thread when no PidControl
        // reads the flow of a PID-controlled node (then nothing depends on the order between the
        // controllers and the controlled links) and every target has exactly one controller.
        std::vector<char> is_target((size_t)std::max(n.n_state, 1), 0);
        for (int s_ : h->list_phase[CONTROL_PID]) is_target[s_] = 1;
        bool ok = (int)h->list_phase[CONTROL_PID].size() == n.n_pid;
        auto it_t = h->hi.find("pid_term_state"), it_s = h->hi.find("pid_target_state");
        auto it_p = h->hi.find("pid_term_ptr");
        if (it_s == h->hi.end() || (int)it_s->second.size() < n.n_pid) ok = false;
        // a controller may read the (still zero) flow of its own target: it does so before it
        // formulates it; the flow of another controller's target must not be read
        if (ok && it_t != h->hi.end() && it_p != h->hi.end() && (int)it_p->second.size() > n.n_pid)
            for (int i = 0; i < n.n_pid; i++)
                for (int k = it_p->second[i]; k < it_p->second[i + 1]; k++) {
                    int s_ = it_t->second[k];
                    if (s_ >= 0 && s_ < n.n_state && is_target[s_] && s_ != it_s->second[i]) ok = false;
                }
        else if (n.n_pid > 0) ok = false;
        else {
            std::vector<char> seen((size_t)std::max(n.n_state, 1), 0);
            for (int i = 0; i < n.n_pid; i++) {
                int s_ = it_s->second[i];
                if (s_ < 0 || s_ >= n.n_state || !is_target[s_] || seen[s_]) ok = false; else seen[s_] = 1;
            }
        }
        h->pid_fused = ok && n.n_pid > 0;
        auto itk = b->d.find("pid_cs_derivative");
        if (itk != b->d.end()) for (double v : itk->second) h->pid_cs_kd_nonzero |= v != 0.0;
        auto itd = b->d.find("pid_derivative");
        h->pid_kd_nonzero = h->pid_cs_kd_nonzero;
        if (itd != b->d.end()) for (double v : itd->second) h->pid_kd_nonzero |= v != 0.0;
        {
            const int want = getenv("RBS_ITERS_PER_PASS") ? atoi(getenv("RBS_ITERS_PER_PASS")) : geti(b, "iters_per_pass", -1);
            h->iters_per_pass = want > 0 ? std::min(want, 10) : 1;
        }
        h->host_threads = getenv("RBS_HOST_THREADS") ? atoi(getenv("RBS_HOST_THREADS")) != 0 : geti(b, "host_threads", 0) != 0;
        h->fuse_small = getenv("RBS_FUSE") ? atoi(getenv("RBS_FUSE")) != 0 : geti(b, "fuse_small", 1) != 0;
    }
    for (int ph = 1; ph <= 2; ph++) {
        std::string key = std::string("list_phase") + char('0' + ph);
        if (upload(h, key, std::vector<int32_t>(h->list_phase[ph].begin(), h->list_phase[ph].end()), true)) return bail("list upload");
        h->d_list_phase[ph] = sp<int>(h, key.c_str());
    }
    // execution mode of the sparse LU: member tiles per block when there are enough members to
    // fill the machine, one launch per level otherwise
    int tile = geti(b, "lu_tile", 0);
    if (tile <= 0) tile = 8;
    h->tile = tile;
    int mode = geti(b, "lu_mode", -1);  // 0 wide, 1 tile, -1 auto
    if (mode < 0) h->tile_mode = (n.B / tile) >= 64 || n.n_lu < 4096;
    else h->tile_mode = mode == 1;
    {
        // ---- shared-memory factor+solve kernel: packed schedule + member tile size
        auto geta = [&](const char* k) -> const std::vector<int32_t>& {
            static const std::vector<int32_t> empty;
            auto it = b->i.find(k);
            return it == b->i.end() ? empty : it->second;
        };
        std::vector<int32_t> blob;
        LinSched& sc = h->sched;
        auto put = [&](const std::vector<int32_t>& v) { int off = (int)blob.size(); blob.insert(blob.end(), v.begin(), v.end()); return off; };
        sc.prod_ptr = put(geta("lu_prod_ptr")); sc.pivot = put(geta("lu_pivot"));
        sc.prod_l = put(geta("lu_prod_l")); sc.prod_u = put(geta("lu_prod_u"));
        sc.diag = put(geta("lu_diag")); sc.perm = put(geta("perm"));
        sc.fwd_row = put(geta("fwd_row")); sc.fwd_ptr = put(geta("fwd_ptr"));
        sc.fwd_l = put(geta("fwd_l")); sc.fwd_k = put(geta("fwd_k"));
        sc.bwd_row = put(geta("bwd_row")); sc.bwd_ptr = put(geta("bwd_ptr"));
        sc.bwd_u = put(geta("bwd_u")); sc.bwd_k = put(geta("bwd_k"));
        sc.lvl_lu = put(geta("lu_level_ptr")); sc.lvl_fwd = put(geta("fwd_level_ptr"));
        sc.lvl_bwd = put(geta("bwd_level_ptr"));
        sc.total = (int)blob.size();
        {
            // per LU slot (LU order): the J entries that land on it, for the L2 tile kernel
            auto &lw = geta("lu_wsrc"), &wp = geta("wsrc_ptr"), &wpos = geta("wsrc_pos"), &wd = geta("wdiag_flag");
            auto itd = b->d.find("wsrc_sign");
            std::vector<int32_t> lptr(1, 0), lcode;
            for (int e = 0; e < n.n_lu; e++) {
                int src = e < (int)lw.size() ? lw[e] : -1;
                if (src >= 0) {
                    for (int k = wp[src]; k < wp[src + 1]; k++)
                        lcode.push_back((wpos[k] << 1) | (itd->second[k] > 0.0 ? 0 : 1));
                    if (wd[src]) lcode.push_back(-1);
                }
                lptr.push_back((int32_t)lcode.size());
            }
            if (upload(h, "lsrc_ptr", lptr, true) || upload(h, "lsrc_code", lcode, true)) return bail("lsrc upload");
            h->d_lsrc_ptr = sp<int>(h, "lsrc_ptr");
            h->d_lsrc_code = sp<int>(h, "lsrc_code");
        }
        sc.n_levels = (int)h->lu_level_ptr.size() - 1;
        sc.n_fwd = (int)h->fwd_level_ptr.size() - 1;
        sc.n_bwd = (int)h->bwd_level_ptr.size() - 1;
        int32_t vmax = 0;
        for (int32_t v : blob) vmax = std::max(vmax, v);
        h->sched16 = vmax < 65535;
        size_t idx_bytes;
        if (h->sched16) {
            std::vector<uint16_t> b16(blob.size());
            for (size_t k = 0; k < blob.size(); k++) b16[k] = blob[k] < 0 ? (uint16_t)0xFFFF : (uint16_t)blob[k];
            idx_bytes = b16.size() * sizeof(uint16_t);
            if (cudaMalloc(&h->d_blob, std::max<size_t>(idx_bytes, 2)) != cudaSuccess) return bail("schedule alloc");
            if (cudaMemcpyAsync(h->d_blob, b16.data(), idx_bytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                cudaStreamSynchronize(h->stream) != cudaSuccess) return bail("schedule upload");
        } else {
            idx_bytes = blob.size() * sizeof(int32_t);
            if (cudaMalloc(&h->d_blob, std::max<size_t>(idx_bytes, 4)) != cudaSuccess) return bail("schedule alloc");
            if (cudaMemcpyAsync(h->d_blob, blob.data(), idx_bytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                cudaStreamSynchronize(h->stream) != cudaSuccess) return bail("schedule upload");
        }
        int want = geti(b, "smem_tile", -1);  // -1 auto (32), 0 off, else 8 / 16 / 32
        int dev_smem = 0;
        cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        int t = want < 0 ? 32 : want;
        if (t != 0 && t != 8 && t != 16 && t != 32) t = 32;
        size_t isz = h->sched16 ? sizeof(uint16_t) : sizeof(int32_t);
        auto bytes_for = [&](int tile) { return (size_t)n.n_reduced * tile * sizeof(double) + (size_t)sc.total * isz + 16; };
        while (t >= 8 && bytes_for(t) + 2048 > (size_t)dev_smem) t >>= 1;
        if (t < 8) t = 0;  // schedule + solution vectors do not fit: level-per-launch fallback
        h->smem_tile = t;
        h->smem_threads = geti(b, "smem_threads", 1024);
        if (t > 0) {
            int bytes = (int)bytes_for(t);
            cudaError_t e1 = cudaSuccess;
#define SET_ATTR(T, IT) e1 = cudaFuncSetAttribute(k_newton_linear<T, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
            if (h->sched16) { if (t == 32) SET_ATTR(32, uint16_t); else if (t == 16) SET_ATTR(16, uint16_t); else SET_ATTR(8, uint16_t); }
            else { if (t == 32) SET_ATTR(32, int32_t); else if (t == 16) SET_ATTR(16, int32_t); else SET_ATTR(8, int32_t); }
#undef SET_ATTR
            if (e1 != cudaSuccess) { cudaGetLastError(); h->smem_tile = 0; }
            h->smem_bytes = bytes;
        }
        // ---- shared-memory kernel (LDU schedule): second packed blob
        int lin_mode = geti(b, "lin_mode", -1);  // -1 auto, 0 wide, 1 L2 tile kernel, 2 shared-memory kernel
        if (b->i.count("ldu_prod_ptr") && want != 0 && (lin_mode < 0 || lin_mode == 2 || lin_mode == 3)) {
            std::vector<int32_t> blob2;
            LduSched& ds = h->dsched;
            auto put2 = [&](const std::vector<int32_t>& v) { int off = (int)blob2.size(); blob2.insert(blob2.end(), v.begin(), v.end()); return off; };
            ds.prod_ptr = put2(geta("ldu_prod_ptr")); ds.flag = put2(geta("ldu_flag"));
            ds.prod_l = put2(geta("ldu_prod_l")); ds.prod_u = put2(geta("ldu_prod_u"));
            ds.prod_p = put2(geta("ldu_prod_p")); ds.diag = put2(geta("ldu_diag")); ds.perm = put2(geta("perm"));
            ds.fwd_row = put2(geta("ldu_fwd_row")); ds.fwd_ptr = put2(geta("ldu_fwd_ptr"));
            ds.fwd_l = put2(geta("ldu_fwd_l")); ds.fwd_k = put2(geta("ldu_fwd_k"));
            ds.bwd_row = put2(geta("ldu_bwd_row")); ds.bwd_ptr = put2(geta("ldu_bwd_ptr"));
            ds.bwd_u = put2(geta("ldu_bwd_u")); ds.bwd_k = put2(geta("ldu_bwd_k"));
            ds.lvl = put2(geta("ldu_level_ptr")); ds.lvl_fwd = put2(geta("ldu_fwd_level_ptr"));
            ds.lvl_bwd = put2(geta("ldu_bwd_level_ptr"));
            {
                // per level: split the product lists over 4 lanes (warp per slot) when that shortens
                // the level's critical path, else one thread per (slot, member)
                auto wide_of = [&](const std::vector<int32_t>& lvlp, const std::vector<int32_t>& ptr) {
                    std::vector<int32_t> w;
                    for (size_t l = 0; l + 1 < lvlp.size(); l++) {
                        int cnt = lvlp[l + 1] - lvlp[l], mx = 0;
                        for (int e = lvlp[l]; e < lvlp[l + 1]; e++) mx = std::max(mx, ptr[e + 1] - ptr[e]);
                        int thin = (cnt + 127) / 128 * (2 + mx), wide = (cnt + 31) / 32 * (3 + (mx + 3) / 4);
                        w.push_back(wide < thin ? 1 : 0);
                    }
                    if (w.empty()) w.push_back(0);
                    return w;
                };
                ds.wide = put2(wide_of(geta("ldu_level_ptr"), geta("ldu_prod_ptr")));
                ds.wide_fwd = put2(wide_of(geta("ldu_fwd_level_ptr"), geta("ldu_fwd_ptr")));
                ds.wide_bwd = put2(wide_of(geta("ldu_bwd_level_ptr"), geta("ldu_bwd_ptr")));
            }
            ds.n_levels = (int)geta("ldu_level_ptr").size() - 1;
            ds.n_fwd = (int)geta("ldu_fwd_level_ptr").size() - 1;
            ds.n_bwd = (int)geta("ldu_bwd_level_ptr").size() - 1;
            {
                auto &lw = geta("ldu_wsrc"), &wp = geta("wsrc_ptr"), &wpos = geta("wsrc_pos"), &wd = geta("wdiag_flag");
                auto itd = b->d.find("wsrc_sign");
                std::vector<int32_t> lptr(1, 0), lcode;
                for (int e = 0; e < n.n_lu; e++) {
                    int src = e < (int)lw.size() ? lw[e] : -1;
                    if (src >= 0) {
                        for (int k = wp[src]; k < wp[src + 1]; k++)
                            lcode.push_back((wpos[k] << 1) | (itd->second[k] > 0.0 ? 0 : 1));
                        if (wd[src]) lcode.push_back(-1);
                    }
                    lptr.push_back((int32_t)lcode.size());
                }
                if (upload(h, "dsrc_ptr", lptr, true) || upload(h, "dsrc_code", lcode, true)) return bail("dsrc upload");
            }
            ds.total = (int)blob2.size();
            int32_t vmax2 = 0;
            for (int32_t v : blob2) vmax2 = std::max(vmax2, v);
            bool s16 = vmax2 < 65535;
            size_t isz2 = s16 ? sizeof(uint16_t) : sizeof(int32_t);
            size_t sb = (size_t)(n.n_lu + n.n_reduced) * 8 * sizeof(double) + (size_t)ds.total * isz2 + 16;
            if (sb + 1024 <= (size_t)dev_smem) {
                if (cudaMalloc(&h->d_blob2, std::max<size_t>(blob2.size() * isz2, 4)) != cudaSuccess) return bail("schedule alloc");
                if (s16) {
                    std::vector<uint16_t> b16(blob2.size());
                    for (size_t k = 0; k < blob2.size(); k++) b16[k] = blob2[k] < 0 ? (uint16_t)0xFFFF : (uint16_t)blob2[k];
                    if (cudaMemcpyAsync(h->d_blob2, b16.data(), b16.size() * 2, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                        cudaStreamSynchronize(h->stream) != cudaSuccess) return bail("schedule upload");
                } else {
                    if (cudaMemcpyAsync(h->d_blob2, blob2.data(), blob2.size() * 4, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                        cudaStreamSynchronize(h->stream) != cudaSuccess) return bail("schedule upload");
                }
                cudaError_t e2 = s16
                    ? cudaFuncSetAttribute(k_newton_linear_smem<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb)
                    : cudaFuncSetAttribute(k_newton_linear_smem<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
                if (e2 == cudaSuccess) { h->lin_smem = true; h->lin_smem_bytes = sb; h->lin_smem16 = s16; }
                else cudaGetLastError();
            }
        }
        // ---- entry schedule kernel (preferred when the tile's slots + schedule fit)
        if (b->i.count("ent_e") && want != 0 && (lin_mode < 0 || lin_mode == 3)) {
            EntSched& es = h->esched;
            es.n_elim = geti(b, "ent_n_elim"); es.n_bwd = geti(b, "ent_n_bwd");
            es.n_slots = n.n_lu + n.n_reduced; es.y0 = n.n_lu;
            // device form: 64-bit level / entry / product words, then perm (uint16)
            auto &ee = geta("ent_e"), &eq = geta("ent_q0"), &ec = geta("ent_cf");
            auto &pl = geta("ent_prod_l"), &pp = geta("ent_prod_p"), &pu = geta("ent_prod_u");
            auto &lp = geta("ent_lvl_ptr"), &tm_ = geta("ent_team_log"), &pm = geta("perm");
            bool ok = n.n_lu + n.n_reduced < 65536 && pl.size() < 65536 && ee.size() < 65536 &&
                      (int)lp.size() == es.n_elim + es.n_bwd + 1 && tm_.size() + 1 == lp.size() &&
                      ee.size() == eq.size() && ee.size() == ec.size() && pl.size() == pp.size() &&
                      pl.size() == pu.size();
            std::vector<unsigned long long> hb;
            size_t o_lvl = 0, o_ent = 0, o_prod = 0, o_perm = 0;
            if (ok) {
                for (size_t k = 0; k + 1 < lp.size(); k++)
                    hb.push_back((unsigned long long)(uint16_t)lp[k] | ((unsigned long long)(uint16_t)lp[k + 1] << 16) |
                                 ((unsigned long long)(uint16_t)tm_[k] << 32));
                o_ent = hb.size();
                for (size_t k = 0; k < ee.size(); k++)
                    hb.push_back((unsigned long long)(uint16_t)ee[k] | ((unsigned long long)(uint16_t)eq[k] << 16) |
                                 ((unsigned long long)(uint16_t)ec[k] << 32));
                o_prod = hb.size();
                for (size_t k = 0; k < pl.size(); k++)
                    hb.push_back((unsigned long long)(uint16_t)pl[k] | ((unsigned long long)(uint16_t)pp[k] << 16) |
                                 ((unsigned long long)(uint16_t)pu[k] << 32));
                o_perm = hb.size();
                for (size_t k = 0; k < pm.size(); k += 4) {
                    unsigned long long w = 0;
                    for (size_t q = 0; q < 4 && k + q < pm.size(); q++) w |= (unsigned long long)(uint16_t)pm[k + q] << (16 * q);
                    hb.push_back(w);
                }
                hb.push_back(0);
            }
            // member tile: small tiles let several blocks share an SM (the level loop is latency
            // bound); the largest of 2 / 4 / 8 members that still leaves room for 4 blocks per SM,
            // else whatever fits
            int et = geti(b, "ent_tile", 0);
            if (et != 2 && et != 4 && et != 8) {
                et = 0;
                for (int cand : {2, 4, 8})
                    if (((size_t)es.n_slots * cand * sizeof(double) + 1024) * 4 <= (size_t)228 * 1024) et = cand;
                if (et == 0) et = 2;
            }
            es.n_words = (int)o_perm;
            // small ensembles (few tiles per SM, the solve's level chain on the critical path of every
            // pass): the schedule words in shared memory behind the slots shorten a level (measured:
            // +8 % at 512 members, +3 % at 1024, -5 % at 2048 where resident tiles matter more)
            auto auto_sched = [&](size_t slot_bytes, size_t sched_bytes) {
                return n_members <= 1024 && (slot_bytes + sched_bytes + 1024) * 2 <= (size_t)227 * 1024 ? 1 : 0;
            };
            es.smem_sched = getenv("RBS_ENT_SMEM_SCHED") ? atoi(getenv("RBS_ENT_SMEM_SCHED"))
                                                         : auto_sched((size_t)es.n_slots * et * sizeof(double), o_perm * 8);
            size_t sb = (size_t)es.n_slots * et * sizeof(double) +
                        (es.smem_sched == 3 ? (o_perm - o_prod) * 8 : es.smem_sched == 2 ? o_prod * 8 : es.smem_sched ? (size_t)es.n_words * 8 : 0);
            if (ok && sb + 256 <= (size_t)dev_smem) {
                if (cudaMalloc(&h->d_ent_blob, hb.size() * 8) != cudaSuccess) return bail("schedule alloc");
                if (cudaMemcpyAsync(h->d_ent_blob, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                    cudaStreamSynchronize(h->stream) != cudaSuccess) return bail("schedule upload");
                const unsigned long long* d = (const unsigned long long*)h->d_ent_blob;
                es.lvl = d + o_lvl; es.ent = d + o_ent; es.prod = d + o_prod;
                es.perm = (const uint16_t*)(d + o_perm);
                cudaError_t e3 =
                    et == 2 ? cudaFuncSetAttribute(k_newton_linear_ent<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb)
                    : et == 4 ? cudaFuncSetAttribute(k_newton_linear_ent<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb)
                              : cudaFuncSetAttribute(k_newton_linear_ent<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
                if (e3 == cudaSuccess) { h->lin_ent = true; h->ent_smem_bytes = sb; h->ent_tile = et; }
                else cudaGetLastError();
            }
        }
        // ---- the same schedule with compacted slots for the full-Newton path
        if (h->lin_ent && b->i.count("entc_e") && !(getenv("RBS_ENT_COMPACT") && atoi(getenv("RBS_ENT_COMPACT")) == 0)) {
            EntSched es = h->esched;
            es.n_slots = geti(b, "entc_n_slots"); es.y0 = geti(b, "entc_y0");
            auto &ee = geta("entc_e"), &eq = geta("entc_q0"), &ec = geta("entc_cf");
            auto &pl = geta("entc_prod_l"), &pp = geta("entc_prod_p"), &pu = geta("entc_prod_u");
            auto &lp = geta("entc_lvl_ptr"), &tm_ = geta("entc_team_log"), &pm = geta("perm");
            auto &le = geta("entc_load_e"), &ls = geta("entc_load_s");
            bool ok = es.n_slots > 0 && es.n_slots <= h->esched.n_slots && es.y0 + n.n_reduced == es.n_slots &&
                      pl.size() < 65536 && ee.size() < 65536 &&
                      (int)lp.size() == es.n_elim + es.n_bwd + 1 && tm_.size() + 1 == lp.size() &&
                      ee.size() == eq.size() && ee.size() == ec.size() && pl.size() == pp.size() &&
                      pl.size() == pu.size() && le.size() == ls.size() && !le.empty();
            if (ok) {
                std::vector<unsigned long long> hb;
                for (size_t k = 0; k + 1 < lp.size(); k++)
                    hb.push_back((unsigned long long)(uint16_t)lp[k] | ((unsigned long long)(uint16_t)lp[k + 1] << 16) |
                                 ((unsigned long long)(uint16_t)tm_[k] << 32));
                size_t o_ent = hb.size();
                for (size_t k = 0; k < ee.size(); k++)
                    hb.push_back((unsigned long long)(uint16_t)ee[k] | ((unsigned long long)(uint16_t)eq[k] << 16) |
                                 ((unsigned long long)(uint16_t)ec[k] << 32));
                size_t o_prod = hb.size();
                for (size_t k = 0; k < pl.size(); k++)
                    hb.push_back((unsigned long long)(uint16_t)pl[k] | ((unsigned long long)(uint16_t)pp[k] << 16) |
                                 ((unsigned long long)(uint16_t)pu[k] << 32));
                size_t o_perm = hb.size();
                for (size_t k = 0; k < pm.size(); k += 4) {
                    unsigned long long w = 0;
                    for (size_t q = 0; q < 4 && k + q < pm.size(); q++) w |= (unsigned long long)(uint16_t)pm[k + q] << (16 * q);
                    hb.push_back(w);
                }
                hb.push_back(0);
                size_t o_load = hb.size();
                for (size_t k = 0; k < le.size(); k += 2) {
                    unsigned long long w = 0;
                    for (size_t q = 0; q < 2 && k + q < le.size(); q++)
                        w |= (unsigned long long)((unsigned)(uint16_t)le[k + q] | ((unsigned)(uint16_t)ls[k + q] << 16)) << (32 * q);
                    hb.push_back(w);
                }
                es.n_words = (int)o_perm;
                es.smem_sched = getenv("RBS_ENT_SMEM_SCHED") ? atoi(getenv("RBS_ENT_SMEM_SCHED"))
                    : (n_members <= 1024 && ((size_t)es.n_slots * h->ent_tile * sizeof(double) + o_perm * 8 + 1024) * 2 <= (size_t)227 * 1024 ? 1 : 0);
                es.n_load = (int)le.size();
                if (cudaMalloc(&h->d_entc_blob, hb.size() * 8) != cudaSuccess) return bail("schedule alloc");
                if (cudaMemcpyAsync(h->d_entc_blob, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                    cudaStreamSynchronize(h->stream) != cudaSuccess) return bail("schedule upload");
                const unsigned long long* d = (const unsigned long long*)h->d_entc_blob;
                es.lvl = d; es.ent = d + o_ent; es.prod = d + o_prod;
                es.perm = (const uint16_t*)(d + o_perm);
                es.load = (const unsigned*)(d + o_load);
                h->esched_c = es;
                h->entc_smem_bytes = (size_t)es.n_slots * h->ent_tile * sizeof(double) +
                                     (es.smem_sched == 3 ? (o_perm - o_prod) * 8 : es.smem_sched == 2 ? o_prod * 8 : es.smem_sched ? (size_t)es.n_words * 8 : 0);
                if (h->entc_smem_bytes > h->ent_smem_bytes) {
                    // the opt-in size of the kernel was set for the plain form
                    cudaError_t e5 = h->ent_tile == 2 ? cudaFuncSetAttribute(k_newton_linear_ent<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->entc_smem_bytes)
                        : h->ent_tile == 4 ? cudaFuncSetAttribute(k_newton_linear_ent<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->entc_smem_bytes)
                                           : cudaFuncSetAttribute(k_newton_linear_ent<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->entc_smem_bytes);
                    if (e5 != cudaSuccess) { cudaGetLastError(); es.smem_sched = 0; h->esched_c = es;
                        h->entc_smem_bytes = (size_t)es.n_slots * h->ent_tile * sizeof(double); }
                }
                h->lin_entc = true;
            }
        }
        if (lin_mode == 0) { h->smem_tile = 0; h->lin_smem = false; h->lin_ent = false; h->lin_entc = false; }
        // ---- M members per thread (full-Newton path, compact slots): the widest tile that fits
        if (h->lin_entc) {
            int want_m = getenv("RBS_LIN_M") ? atoi(getenv("RBS_LIN_M")) : geti(b, "lin_m", 0);
            int want_sched = getenv("RBS_LIN_M_SCHED") ? atoi(getenv("RBS_LIN_M_SCHED")) : 1;
            for (int cand : {8, 4}) {
                if (want_m == 0 || (want_m > 0 && want_m != cand)) continue;
                size_t slots = (size_t)h->esched_c.n_slots * cand * sizeof(double);
                size_t sched = (size_t)h->esched_c.n_words * 8;
                int use_sched = want_sched && slots + sched + 256 <= (size_t)dev_smem;
                size_t sb = slots + (use_sched ? sched : 0);
                if (sb + 256 > (size_t)dev_smem) continue;
                cudaError_t e6 = cand == 8
                    ? cudaFuncSetAttribute(k_newton_linear_m<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb)
                    : cudaFuncSetAttribute(k_newton_linear_m<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
                if (e6 != cudaSuccess) { cudaGetLastError(); continue; }
                h->lin_m = cand; h->lin_m_sched = use_sched; h->lin_m_smem = sb;
                break;
            }
        }
        // ---- member per lane on global slots (rbs_lane.cuh): the plain entry schedule unrolled into
        // one linear tape of 16-byte records per warp, offsets ready-made for this ensemble size
        if (h->lin_ent) {
            int wl = getenv("RBS_LIN_LANE") ? atoi(getenv("RBS_LIN_LANE")) : geti(b, "lin_lane", 32);
            h->lane_min = getenv("RBS_LIN_LANE_MIN") ? atoi(getenv("RBS_LIN_LANE_MIN")) : geti(b, "lin_lane_min", 2048);
            const int NW = wl <= 0 ? 0 : wl >= 32 ? 32 : wl >= 16 ? 16 : 8;
            auto &ee = geta("ent_e"), &eq = geta("ent_q0"), &ec = geta("ent_cf");
            auto &pl = geta("ent_prod_l"), &pp = geta("ent_prod_p"), &pu = geta("ent_prod_u");
            auto &lp = geta("ent_lvl_ptr"), &tm_ = geta("ent_team_log"), &pm = geta("perm");
            const int y0 = n.n_lu, n_lv = h->esched.n_elim + h->esched.n_bwd;
            const unsigned long long Bq = (unsigned long long)n.B;
            bool ok = NW > 0 && (unsigned long long)(n.n_lu + n.n_reduced) * Bq < (1ull << 32);
            std::map<int, int> piv_of;
            std::vector<unsigned> piv_off;
            if (ok)
                for (size_t k = 0; k < ee.size(); k++) {
                    if (ec[k] & 0x400) ok = false;  // FRESH belongs to the compact layout
                    if (ec[k] & 0x100) { piv_of[ee[k]] = (int)piv_off.size(); piv_off.push_back((unsigned)(ee[k] * Bq)); }
                }
            auto pidx = [&](int slot) -> unsigned {
                auto it = piv_of.find(slot);
                if (it == piv_of.end()) { ok = false; return 0u; }
                return (unsigned)it->second * 256u;
            };
            std::vector<std::vector<uint4>> tp(std::max(NW, 1));
            auto kind_of = [&](int k, bool bwd) -> unsigned {
                const int e = ee[k];
                const bool rcp = (ec[k] & 0x100) != 0, scale = (ec[k] & 0x200) != 0;
                if (bwd) { if (e < y0 || !scale || rcp) ok = false; return RBS_LANE_K_YB; }
                if (scale || (rcp && e >= y0)) ok = false;
                return rcp ? RBS_LANE_K_RCP : e >= y0 ? RBS_LANE_K_YF : RBS_LANE_K_FAC;
            };
            auto entry_hdr = [&](int k, int cnt, bool bwd) {
                const int e = ee[k];
                const unsigned kind = kind_of(k, bwd);
                unsigned py = 0, xo = 0;
                if (kind == RBS_LANE_K_RCP) py = pidx(e);
                if (kind == RBS_LANE_K_YB) {
                    py = pidx((ec[k] & 0xFF) ? pp[eq[k]] : eq[k]);
                    xo = (unsigned)(pm[e - y0] * Bq);
                }
                return make_uint4((unsigned)(e * Bq), py, (unsigned)cnt | (kind << 16), xo);
            };
            auto product = [&](int q) { return make_uint4((unsigned)(pl[q] * Bq), pidx(pp[q]), (unsigned)(pu[q] * Bq), 0u); };
            for (int lv = 0; ok && lv < n_lv; lv++) {
                const int e0 = lp[lv], e1 = lp[lv + 1], tl = tm_[lv], ts = 1 << tl;
                const bool bwd = lv >= h->esched.n_elim;
                if (tl > 0 && ((e1 - e0) << tl) <= NW) {
                    for (int w = 0; w < NW; w++) {
                        const bool has = w < ((e1 - e0) << tl);
                        const int t = w & (ts - 1);
                        tp[w].push_back(make_uint4(has ? 1u : 0u, 1u, (unsigned)tl, (unsigned)t));
                        if (!has) continue;
                        const int k = e0 + (w >> tl), q0 = eq[k], cnt = ec[k] & 0xFF;
                        std::vector<int> qs;
                        for (int q = q0 + t; q < q0 + cnt; q += ts) qs.push_back(q);
                        tp[w].push_back(entry_hdr(k, (int)qs.size(), bwd));
                        for (int q : qs) tp[w].push_back(product(q));
                    }
                } else {
                    // entries to warps: longest first to the least loaded warp (instruction estimate);
                    // a warp's entries then go into runs of one kind and, without teams, one count
                    std::vector<int> ks;
                    for (int k = e0; k < e1; k++) ks.push_back(k);
                    auto cost = [&](int k) { return 30 + 11 * (ec[k] & 0xFF) + ((ec[k] & 0x100) ? 25 : 0); };
                    std::stable_sort(ks.begin(), ks.end(), [&](int a, int b2) { return cost(a) > cost(b2); });
                    std::vector<std::vector<int>> mine(NW);
                    std::vector<long long> load(NW, 0);
                    for (int k : ks) {
                        int best = 0;
                        for (int w = 1; w < NW; w++) if (load[w] < load[best]) best = w;
                        mine[best].push_back(k);
                        load[best] += cost(k);
                    }
                    for (int w = 0; w < NW; w++) {
                        auto code = [&](int k) -> unsigned { const unsigned cnt = ec[k] & 0xFF; return tl == 0 && cnt <= 4 ? cnt : RBS_LANE_DYN; };
                        auto key = [&](int k) { return (kind_of(k, bwd) << 8) | code(k); };
                        std::stable_sort(mine[w].begin(), mine[w].end(), [&](int a, int b2) { return key(a) < key(b2); });
                        std::vector<uint4> body;
                        unsigned n_runs = 0;
                        for (size_t a = 0; a < mine[w].size();) {
                            size_t z = a;
                            while (z < mine[w].size() && key(mine[w][z]) == key(mine[w][a])) z++;
                            body.push_back(make_uint4(kind_of(mine[w][a], bwd), code(mine[w][a]), (unsigned)(z - a), 0u));
                            for (size_t y = a; y < z; y++) {
                                const int k = mine[w][y], q0 = eq[k], cnt = ec[k] & 0xFF;
                                body.push_back(entry_hdr(k, cnt, bwd));
                                for (int q = q0; q < q0 + cnt; q++) body.push_back(product(q));
                            }
                            n_runs++;
                            a = z;
                        }
                        tp[w].push_back(make_uint4(n_runs, 0u, (unsigned)tl, 0u));
                        tp[w].insert(tp[w].end(), body.begin(), body.end());
                    }
                }
            }
            const size_t sb = ((size_t)NW * 32 + piv_off.size() * 32) * sizeof(double);
            if (ok && sb + 1024 <= (size_t)dev_smem) {
                std::vector<uint4> recs;
                std::vector<int> start(NW);
                for (int w = 0; w < NW; w++) { start[w] = (int)recs.size(); recs.insert(recs.end(), tp[w].begin(), tp[w].end()); }
                recs.resize(recs.size() + 32, make_uint4(0, 0, 0, 0));  // the reader is one record and two lines ahead
                std::vector<int> iperm(n.n_reduced, 0);
                for (int i = 0; i < n.n_reduced && i < (int)pm.size(); i++) iperm[pm[i]] = i;
                const size_t b_rec = recs.size() * sizeof(uint4), b_start = (size_t)NW * sizeof(int), b_piv = (piv_off.size() + 3) / 4 * 4 * sizeof(unsigned);
                const size_t b_ip = iperm.size() * sizeof(int);
                if (cudaMalloc(&h->d_lane_blob, b_rec + b_start + b_piv + b_ip + 16) != cudaSuccess) return bail("lane tape alloc");
                char* d = (char*)h->d_lane_blob;
                if (cudaMemcpyAsync(d, recs.data(), b_rec, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                    cudaMemcpyAsync(d + b_rec, start.data(), b_start, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                    cudaMemcpyAsync(d + b_rec + b_start, piv_off.data(), piv_off.size() * sizeof(unsigned), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                    cudaMemcpyAsync(d + b_rec + b_start + b_piv, iperm.data(), b_ip, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                    cudaStreamSynchronize(h->stream) != cudaSuccess) return bail("lane tape upload");
                LaneTape& lt = h->lane_tape;
                lt.rec = (const uint4*)d; lt.start = (const int*)(d + b_rec); lt.piv_off = (const unsigned*)(d + b_rec + b_start);
                lt.n_lv = n_lv; lt.n_piv = (int)piv_off.size(); lt.iperm = (const int*)(d + b_rec + b_start + b_piv);
                cudaError_t e7 = NW == 32 ? cudaFuncSetAttribute(k_newton_linear_lane<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb)
                               : NW == 16 ? cudaFuncSetAttribute(k_newton_linear_lane<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb)
                                          : cudaFuncSetAttribute(k_newton_linear_lane<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
                if (e7 == cudaSuccess) { h->lin_lane = NW; h->lane_smem = sb; }
                else cudaGetLastError();
            }
        }
        apply_carveout(h);
        // ---- cooperative window kernel: needs the entry schedule with 2-member tiles in shared
        // memory and a grid that is co-resident
        if (h->lin_ent) {
            size_t sb2 = (size_t)h->esched.n_slots * 2 * sizeof(double);
            int coop_ok = 0, per_sm = 0, n_sm = 0;
            cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, device);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
            if (coop_ok && sb2 + 8192 <= (size_t)dev_smem &&
                cudaFuncSetAttribute(k_window_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb2) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_window_coop, RBS_COOP_THREADS, sb2) == cudaSuccess &&
                per_sm > 0) {
                h->coop_blocks = per_sm * n_sm;
                if (cudaMalloc(&h->d_coop_counters, 4 * sizeof(int)) != cudaSuccess ||
                    cudaMallocHost(&h->h_coop_counters, 4 * sizeof(int)) != cudaSuccess) return bail("coop alloc");
                int cm = geti(b, "coop_members", -1);
                h->coop_members = cm >= 0 ? cm : 32;  // measured: pays below a few dozen members
            } else cudaGetLastError();
            // tile window kernel: a block per member tile runs the whole window (rbs_tile.cuh)
            int tw = geti(b, "tile_window", -1);
            if (tw < 0) tw = 0;
            if (tw == 2 || tw == 4) {
                size_t words = std::max<size_t>((size_t)h->esched.n_slots, (size_t)h->n_part * RBS_DZ_ROWS);
                size_t sbt = words * tw * sizeof(double);
                cudaError_t e4 = sbt + 1024 > (size_t)dev_smem ? cudaErrorInvalidValue
                    : tw == 2 ? cudaFuncSetAttribute(k_window_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbt)
                              : cudaFuncSetAttribute(k_window_tile<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbt);
                if (e4 == cudaSuccess) {
                    h->tile_window = tw; h->tile_smem = sbt;
                    if (!h->d_coop_counters &&
                        (cudaMalloc(&h->d_coop_counters, 4 * sizeof(int)) != cudaSuccess ||
                         cudaMallocHost(&h->h_coop_counters, 4 * sizeof(int)) != cudaSuccess)) return bail("tile alloc");
                } else cudaGetLastError();
            }
        }
    }
    // programmatic dependent launch pays where a pass is a chain of short kernels (measured: +9 % at
    // 512 members, -4 % at 4096, where early-resident blocks take slots from the other groups)
    {
        const int want_pdl = getenv("RBS_PDL") ? atoi(getenv("RBS_PDL")) : geti(b, "pdl", -1);
        h->pdl = want_pdl < 0 ? n_members <= 1024 : want_pdl != 0;
    }
    h->r2_pass = getenv("RBS_R2_PASS") ? atoi(getenv("RBS_R2_PASS")) : geti(b, "r2_pass", 0);
    h->fin_cluster = getenv("RBS_FIN_CLUSTER") ? atoi(getenv("RBS_FIN_CLUSTER")) : geti(b, "fin_cluster", 4);
    bind_net(h);
    n.cc_tab_offset = geti(b, "cc_tab_offset");
    if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return bail("rbs_create: device initialisation failed");
    return h;
}

void rbs_destroy(rbs_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto& kv : h->shared) if (kv.second.p) cudaFree(kv.second.p);
    for (auto& kv : h->member) if (kv.second.p) cudaFree(kv.second.p);
    if (h->staging) cudaFree(h->staging);
    for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (h->h_active) cudaFreeHost(h->h_active);
    free_groups(h);
    for (auto& kv : h->shadow_in) if (kv.second) cudaFree(kv.second);
    for (auto& kv : h->shadow_out) if (kv.second) cudaFree(kv.second);
    for (int k = 0; k < 6; k++) { if (h->fw_live[k]) cudaFree(h->fw_live[k]); if (h->fw_shadow[k]) cudaFree(h->fw_shadow[k]); }
    for (int k = 0; k < 2; k++) {
        if (h->fw_res[k]) cudaFree(h->fw_res[k]);
        if (h->fw_flow[k]) cudaFree(h->fw_flow[k]);
        if (h->ev_d2h[k]) cudaEventDestroy(h->ev_d2h[k]);
    }
    if (h->ev_h2d) cudaEventDestroy(h->ev_h2d);
    if (h->ev_win) cudaEventDestroy(h->ev_win);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->ev_main) cudaEventDestroy(h->ev_main);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d_blob) cudaFree(h->d_blob);
    if (h->d_blob2) cudaFree(h->d_blob2);
    if (h->d_ent_blob) cudaFree(h->d_ent_blob);
    if (h->d_entc_blob) cudaFree(h->d_entc_blob);
    if (h->d_lane_blob) cudaFree(h->d_lane_blob);
    if (h->d_coop_counters) cudaFree(h->d_coop_counters);
    if (h->h_coop_counters) cudaFreeHost(h->h_coop_counters);
    if (h->d_alist) cudaFree(h->d_alist);
    if (h->d_llist) cudaFree(h->d_llist);
    if (h->d_counts) cudaFree(h->d_counts);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int64_t rbs_kernel_launches(const rbs_handle* h) { return h ? h->launches : 0; }

int rbs_set_param_f64(rbs_handle* h, const char* key, const double* host, int64_t n) {
    if (!h || !key) return fail("rbs_set_param_f64: null argument");
    cudaSetDevice(h->device);
    auto it = h->shared.find(key);
    if (it == h->shared.end() || it->second.is_int) return fail(std::string("unknown f64 parameter ") + key);
    if ((size_t)std::max<int64_t>(n, 1) * sizeof(double) != it->second.bytes) return fail(std::string("parameter size mismatch: ") + key);
    if (n > 0) CUDA_TRY(cudaMemcpyAsync(it->second.p, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));  // host buffer is borrowed only for the call
    if (!strcmp(key, "pid_derivative")) {
        bool nz = h->pid_cs_kd_nonzero;
        for (int64_t k = 0; k < n; k++) nz |= host[k] != 0.0;
        h->pid_kd_nonzero = nz;
    }
    return 0;
}
int rbs_set_param_i32(rbs_handle* h, const char* key, const int32_t* host, int64_t n) {
    if (!h || !key) return fail("rbs_set_param_i32: null argument");
    cudaSetDevice(h->device);
    auto it = h->shared.find(key);
    if (it == h->shared.end() || !it->second.is_int) return fail(std::string("unknown i32 parameter ") + key);
    if ((size_t)std::max<int64_t>(n, 1) * sizeof(int32_t) != it->second.bytes) return fail(std::string("parameter size mismatch: ") + key);
    if (n > 0) CUDA_TRY(cudaMemcpyAsync(it->second.p, host, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

int rbs_set_member_param_f64(rbs_handle* h, const char* key, const double* host, int64_t n_entity) {
    if (!h || !key) return fail("rbs_set_member_param_f64: null argument");
    cudaSetDevice(h->device);
    Net& n = h->net;
    struct PK { const char* key; int count; };
    const PK keys[] = {{"mn_manning_n", n.n_manning}, {"lr_resistance", n.n_lr},
                       {"pump_flow_rate", n.n_pump}, {"outlet_flow_rate", n.n_outlet},
                       {"lb_level", n.n_lb}};
    const PK* pk = nullptr;
    for (auto& k : keys) if (!strcmp(key, k.key)) pk = &k;
    if (!pk) return fail(std::string("rbs_set_member_param_f64: no per-member override for ") + key);
    std::string mk = std::string("pm_") + key;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!host) {  // back to the shared value
        auto it = h->member.find(mk);
        if (it != h->member.end()) { cudaFree(it->second.p); h->member.erase(it); h->member_entities.erase(mk); }
        bind_net(h);
        return 0;
    }
    if (n_entity != pk->count) return fail(std::string("parameter size mismatch: ") + key);
    if (n_entity == 0) return 0;
    if (h->member.find(mk) == h->member.end() && alloc_member(h, mk, n_entity)) return 1;
    CUDA_TRY(cudaMemcpyAsync(h->member[mk].p, host, (size_t)n_entity * h->B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    bind_net(h);
    return check_async(h, "rbs_set_member_param_f64");
}

int rbs_set_member_f64(rbs_handle* h, const char* key, const double* host, int64_t n_entity, int32_t broadcast) {
    if (!h || !key) return fail("rbs_set_member_f64: null argument");
    cudaSetDevice(h->device);
    auto it = h->member.find(key);
    if (it == h->member.end() || it->second.is_int) return fail(std::string("unknown member array ") + key);
    if (h->member_entities[key] != n_entity) return fail(std::string("member array size mismatch: ") + key);
    if (n_entity == 0) return 0;
    if (broadcast) {
        if (ensure_staging(h, (size_t)n_entity * sizeof(double))) return 1;
        CUDA_TRY(cudaMemcpyAsync(h->staging, host, (size_t)n_entity * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        k_broadcast<<<grid_for(n_entity * (long long)h->B), 256, 0, h->stream>>>((double*)it->second.p, h->staging, (int)n_entity, h->B);
        h->launches++;
    } else {
        CUDA_TRY(cudaMemcpyAsync(it->second.p, host, (size_t)n_entity * h->B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_set_member_f64");
}
// make staged uploads live: swap shadow and live buffers once the copy stream has delivered them
static int commit_staged(rbs_handle* h) {
    if (h->staged.empty()) return 0;
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
    for (auto& key : h->staged) std::swap(h->member[key].p, h->shadow_in[key]);
    h->staged.clear();
    bind_net(h);
    return 0;
}
int rbs_set_member_f64_async(rbs_handle* h, const char* key, const double* pinned_host, int64_t n_entity) {
    if (!h || !key || !pinned_host) return fail("rbs_set_member_f64_async: null argument");
    cudaSetDevice(h->device);
    auto it = h->member.find(key);
    if (it == h->member.end() || it->second.is_int) return fail(std::string("unknown member array ") + key);
    if (h->member_entities[key] != n_entity) return fail(std::string("member array size mismatch: ") + key);
    if (n_entity == 0) return 0;
    void*& sh = h->shadow_in[key];
    if (!sh) CUDA_TRY(cudaMalloc(&sh, it->second.bytes));
    for (auto& k : h->staged) if (k == key) return fail(std::string("already staged: ") + key);
    CUDA_TRY(cudaMemcpyAsync(sh, pinned_host, (size_t)n_entity * h->B * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(cudaEventRecord(h->ev_copy, h->copy_stream));
    h->staged.push_back(key);
    return 0;
}
int rbs_get_member_f64_async(rbs_handle* h, const char* key, double* pinned_host, int64_t n_entity) {
    if (!h || !key || !pinned_host) return fail("rbs_get_member_f64_async: null argument");
    cudaSetDevice(h->device);
    auto it = h->member.find(key);
    if (it == h->member.end() || it->second.is_int) return fail(std::string("unknown member array ") + key);
    if (h->member_entities[key] != n_entity) return fail(std::string("member array size mismatch: ") + key);
    if (n_entity == 0) return 0;
    void*& sh = h->shadow_out[key];
    if (!sh) CUDA_TRY(cudaMalloc(&sh, it->second.bytes));
    size_t nb = (size_t)n_entity * h->B * sizeof(double);
    // snapshot on the compute stream (ordered after the step), transfer on the copy stream; the
    // previous transfer out of this snapshot buffer must have left first
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
    CUDA_TRY(cudaMemcpyAsync(sh, it->second.p, nb, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaEventRecord(h->ev_main, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_main, 0));
    CUDA_TRY(cudaMemcpyAsync(pinned_host, sh, nb, cudaMemcpyDeviceToHost, h->copy_stream));
    CUDA_TRY(cudaEventRecord(h->ev_copy, h->copy_stream));
    return 0;
}
int rbs_get_member_f64(rbs_handle* h, const char* key, double* host, int64_t n_entity) {
    if (!h || !key || !host) return fail("rbs_get_member_f64: null argument");
    cudaSetDevice(h->device);
    auto it = h->member.find(key);
    if (it == h->member.end() || it->second.is_int) return fail(std::string("unknown member array ") + key);
    if (h->member_entities[key] != n_entity) return fail(std::string("member array size mismatch: ") + key);
    if (n_entity == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(host, it->second.p, (size_t)n_entity * h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}
int rbs_copy_member_to_device(rbs_handle* h, const char* key, void* dst_device, int64_t n_entity) {
    if (!h || !key || !dst_device) return fail("rbs_copy_member_to_device: null argument");
    cudaSetDevice(h->device);
    auto it = h->member.find(key);
    if (it == h->member.end() || it->second.is_int) return fail(std::string("unknown member array ") + key);
    if (h->member_entities[key] != n_entity) return fail(std::string("member array size mismatch: ") + key);
    if (n_entity == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(dst_device, it->second.p, (size_t)n_entity * h->B * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}
int rbs_get_member_i32(rbs_handle* h, const char* key, int32_t* host, int64_t n_entity) {
    if (!h || !key || !host) return fail("rbs_get_member_i32: null argument");
    cudaSetDevice(h->device);
    auto it = h->member.find(key);
    if (it == h->member.end() || !it->second.is_int) return fail(std::string("unknown int member array ") + key);
    if (h->member_entities[key] != n_entity) return fail(std::string("member array size mismatch: ") + key);
    if (n_entity == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(host, it->second.p, (size_t)n_entity * h->B * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}
int rbs_apply_discrete_control(rbs_handle* h) {
    if (!h) return fail("rbs_apply_discrete_control: null handle");
    cudaSetDevice(h->device);
    Net& n = h->net;
    n.win_t0 = h->tprev;
    if (n.n_dc > 0) LAUNCH(h, k_discrete_control, (long long)n.n_dc * n.B, n, 1);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_apply_discrete_control");
}
int rbs_set_tprev(rbs_handle* h, double tprev) {
    if (!h) return fail("rbs_set_tprev: null handle");
    h->tprev = tprev;
    return 0;
}
int rbs_set_solver(rbs_handle* h, double abstol, double reltol, int32_t max_iter) {
    if (!h) return fail("rbs_set_solver: null handle");
    if (!(abstol >= 0) || !(reltol >= 0) || max_iter < 1) return fail("rbs_set_solver: invalid option");
    h->abstol = abstol; h->reltol = reltol; h->max_iter = max_iter;
    return 0;
}
int rbs_set_stepper(rbs_handle* h, int32_t full_newton, int32_t predictor, int32_t max_halvings,
                    int32_t grow_iters, int32_t early_exit, int32_t shrink_log2) {
    if (!h) return fail("rbs_set_stepper: null handle");
    if (max_halvings < 0 || max_halvings > 60 || grow_iters < 0 || shrink_log2 < 1 || shrink_log2 > 8)
        return fail("rbs_set_stepper: invalid option");
    h->full_newton = full_newton != 0; h->predictor = predictor != 0;
    cudaSetDevice(h->device);
    apply_carveout(h);
    h->max_halvings = max_halvings; h->grow_iters = grow_iters;
    h->early_exit = early_exit != 0; h->shrink_log2 = shrink_log2;
    return 0;
}
int rbs_set_groups(rbs_handle* h, int32_t n_groups) {
    if (!h) return fail("rbs_set_groups: null handle");
    if (n_groups < 0 || n_groups > 64) return fail("rbs_set_groups: invalid group count");
    cudaSetDevice(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return make_groups(h, n_groups);
}
int rbs_get_stats(rbs_handle* h, int64_t* out6) {
    if (!h || !out6) return fail("rbs_get_stats: null argument");
    cudaSetDevice(h->device);
    Net& n = h->net;
    std::vector<int> buf((size_t)n.B * 3);
    CUDA_TRY(cudaMemcpyAsync(buf.data(), n.cnt_iters, n.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(buf.data() + n.B, n.cnt_sub, n.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(buf.data() + 2 * n.B, n.cnt_rej, n.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int64_t it = 0, sub = 0, rej = 0;
    for (int m = 0; m < n.B; m++) { it += buf[m]; sub += buf[n.B + m]; rej += buf[2 * n.B + m]; }
    out6[0] = h->n_steps; out6[1] = h->n_passes; out6[2] = sub; out6[3] = h->n_failed;
    out6[4] = it; out6[5] = rej;
    return 0;
}

int rbs_profile(rbs_handle* h, int32_t enable) {
    if (!h) return fail("rbs_profile: null handle");
    cudaSetDevice(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    h->prof.clear();
    h->profiling = enable != 0;
    return 0;
}
int rbs_profile_report(rbs_handle* h, char* buf, int32_t len) {
    if (!h || !buf || len < 1) return fail("rbs_profile_report: null argument");
    cudaSetDevice(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    std::map<std::string, std::pair<int64_t, double>> agg;
    for (auto& r : h->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& a = agg[r.name];
            a.first++;
            a.second += ms;
        }
    }
    std::string out;
    char line[256];
    for (auto& kv : agg) {
        snprintf(line, sizeof line, "%s\t%lld\t%.6f\n", kv.first.c_str(), (long long)kv.second.first, kv.second.second);
        out += line;
    }
    strncpy(buf, out.c_str(), (size_t)len - 1);
    buf[len - 1] = 0;
    return 0;
}

void* rbs_device_ptr(rbs_handle* h, const char* key) {
    if (!h || !key) return nullptr;
    auto it = h->member.find(key);
    return it == h->member.end() ? nullptr : it->second.p;
}
void* rbs_stream(rbs_handle* h) { return h ? (void*)h->stream : nullptr; }
void* rbs_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0) { fail("rbs_host_alloc: size must be positive"); return nullptr; }
    cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { fail(std::string("rbs_host_alloc: ") + cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    return p;
}
void rbs_host_free(void* p) { if (p) cudaFreeHost(p); }

// The one collective of the design: all-gather of a per-member array over the
// ranks of the caller's NCCL communicator.  NCCL is bound at call time (the library the host
// process already uses, else libnccl.so.2), so that librbs.so itself loads without it.
int rbs_gather(rbs_handle* h, void* nccl_comm, const char* key, int64_t n_entity, void* recv_device) {
    if (!h || !nccl_comm || !key || !recv_device) return fail("rbs_gather: null argument");
    cudaSetDevice(h->device);
    auto it = h->member.find(key);
    if (it == h->member.end() || it->second.is_int) return fail(std::string("unknown member array ") + key);
    if (h->member_entities[key] != n_entity) return fail(std::string("member array size mismatch: ") + key);
    typedef ncclResult_t (*all_gather_fn)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    typedef const char* (*err_fn)(ncclResult_t);
    static all_gather_fn all_gather = nullptr;
    static err_fn err_string = nullptr;
    if (!all_gather) {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return fail(std::string("rbs_gather: NCCL not found: ") + dlerror());
        all_gather = (all_gather_fn)dlsym(lib, "ncclAllGather");
        err_string = (err_fn)dlsym(lib, "ncclGetErrorString");
        if (!all_gather) return fail("rbs_gather: ncclAllGather not found in the NCCL library");
    }
    const size_t count = (size_t)n_entity * h->B;
    if (count == 0) return 0;
    ncclResult_t r = all_gather(it->second.p, recv_device, count, ncclFloat64, (ncclComm_t)nccl_comm, h->stream);
    if (r != ncclSuccess)
        return fail(std::string("rbs_gather: ncclAllGather: ") + (err_string ? err_string(r) : "error"));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_gather");
}

int rbs_synchronize(rbs_handle* h) {
    if (!h) return fail("rbs_synchronize: null handle");
    cudaSetDevice(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
    return check_async(h, "rbs_synchronize");
}

// ------------------------------------------------------------------------------------------ seams
int rbs_reduce(rbs_handle* h, const double* u, double* u_reduced) {
    if (!h || !u || !u_reduced) return fail("rbs_reduce: null argument");
    cudaSetDevice(h->device);
    Net& n = h->net;
    size_t nb = (size_t)n.n_state * n.B * sizeof(double);
    if (ensure_staging(h, nb)) return 1;
    CUDA_TRY(cudaMemcpyAsync(h->staging, u, nb, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(h, k_reduce<0>, (long long)n.n_reduced * n.B, n, h->staging, nullptr, n.ur, 0.0, 0.0);
    CUDA_TRY(cudaMemcpyAsync(u_reduced, n.ur, (size_t)n.n_reduced * n.B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_reduce");
}

static int copy_out(rbs_handle* h, double* host, const double* dev, int64_t n_entity) {
    if (!host || n_entity == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(host, dev, (size_t)n_entity * h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

int rbs_rhs(rbs_handle* h, const double* u_reduced, double t, double* du, double* storage,
            double* level, double* area, double* factor) {
    if (!h || !u_reduced) return fail("rbs_rhs: null argument");
    cudaSetDevice(h->device);
    Net& n = h->net;
    CUDA_TRY(cudaMemcpyAsync(n.ur, u_reduced, (size_t)n.n_reduced * n.B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemsetAsync(n.du, 0, (size_t)n.n_state * n.B * sizeof(double), h->stream));
    LAUNCH(h, k_exact, (long long)n.n_basin * n.B, n, t, h->tprev);
    launch_rhs<false, 0>(h, n.ur);
    if (copy_out(h, du, n.du, n.n_state) || copy_out(h, storage, n.storage, n.n_basin) ||
        copy_out(h, level, n.level, n.n_basin) || copy_out(h, area, n.area, n.n_basin) ||
        copy_out(h, factor, n.factor, n.n_basin)) return 1;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_rhs");
}

int rbs_jac(rbs_handle* h, const double* u_reduced, double t, double gamma, double* j_vals, double* w_vals) {
    if (!h || !u_reduced) return fail("rbs_jac: null argument");
    if (!(gamma > 0)) return fail("rbs_jac: gamma must be positive");
    cudaSetDevice(h->device);
    Net& n = h->net;
    CUDA_TRY(cudaMemcpyAsync(n.ur, u_reduced, (size_t)n.n_reduced * n.B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemsetAsync(n.du, 0, (size_t)n.n_state * n.B * sizeof(double), h->stream));
    CUDA_TRY(cudaMemsetAsync(n.jv, 0, (size_t)std::max(n.nnz_j, 1) * n.B * sizeof(double), h->stream));
    LAUNCH(h, k_exact, (long long)n.n_basin * n.B, n, t, h->tprev);
    launch_rhs<true, 0>(h, n.ur);
    launch_factor<false>(h, 1.0 / gamma);
    if (copy_out(h, j_vals, n.jv, n.nnz_j) || copy_out(h, w_vals, n.wv, n.nnz_w)) return 1;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_jac");
}

int rbs_solve(rbs_handle* h, const double* b, double* x) {
    if (!h || !b || !x) return fail("rbs_solve: null argument");
    if (!h->factor_valid) return fail("rbs_solve: no factorisation; call rbs_jac first");
    cudaSetDevice(h->device);
    Net& n = h->net;
    size_t nb = (size_t)n.n_reduced * n.B * sizeof(double);
    CUDA_TRY(cudaMemcpyAsync(n.br, b, nb, cudaMemcpyHostToDevice, h->stream));
    double* c = mp(h, "cvec");
    launch_solve<false>(h, n.br, c);
    CUDA_TRY(cudaMemcpyAsync(x, c, nb, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_solve");
}

int rbs_limit_flow(rbs_handle* h, double* u, const double* uprev, double dt, double t) {
    if (!h || !u || !uprev) return fail("rbs_limit_flow: null argument");
    cudaSetDevice(h->device);
    Net& n = h->net;
    long long B = n.B;
    size_t nb = (size_t)n.n_state * n.B * sizeof(double);
    CUDA_TRY(cudaMemcpyAsync(n.u, u, nb, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(n.uprev, uprev, nb, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(h, k_exact, (long long)n.n_basin * B, n, t, h->tprev);
    LAUNCH(h, (k_basin<false, 2>), (long long)n.n_reduced * B, n, nullptr, false, Pred{nullptr, 0, nullptr}, 0);
    LAUNCH(h, k_limit_flow, (long long)n.off_integral * B, n, n.u, n.uprev, dt,
           sp<int>(h, "ud_demand_from_timeseries"), sp<double>(h, "ud_alloc_total"));
    CUDA_TRY(cudaMemcpyAsync(u, n.u, nb, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_limit_flow");
}

int rbs_accept_step(rbs_handle* h, double dt) {
    if (!h) return fail("rbs_accept_step: null handle");
    if (!(dt >= 0)) return fail("rbs_accept_step: dt must not be negative");
    cudaSetDevice(h->device);
    Net& n = h->net;
    LAUNCH(h, k_accumulate_forcing, (long long)(n.n_basin + n.n_fb) * n.B, n, dt);
    h->tprev += dt;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_accept_step");
}

int rbs_refresh(rbs_handle* h, double t) {
    if (!h) return fail("rbs_refresh: null handle");
    cudaSetDevice(h->device);
    Net& n = h->net;
    long long B = n.B;
    CUDA_TRY(cudaMemsetAsync(n.du, 0, (size_t)n.n_state * n.B * sizeof(double), h->stream));
    LAUNCH(h, k_exact, (long long)n.n_basin * B, n, t, h->tprev);
    launch_rhs<false, 2>(h, nullptr);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_async(h, "rbs_refresh");
}

// Newton linear solve of the members in the launch domain: fused factor+solve kernel, or the
// level-per-launch kernels when the schedule does not fit in shared memory.
static void launch_linear(rbs_handle* h, double* c, Pred nw) {
    Net& n = net_of(h);
    if (h->lin_lane > 0 && h->lin_ent && n.NA >= h->lane_min) {
        n.lane_iperm = h->lane_tape.iperm;
        if (!getenv("RBS_TIME_SOLVE_ONLY"))
        LAUNCH(h, k_assemble_linear, (long long)(n.n_lu + n.n_reduced) * ((n.NA + 32 * RBS_MPT - 1) / (32 * RBS_MPT)) * 32, n,
               sp<int>(h, "dsrc_ptr"), sp<int>(h, "dsrc_code"));
        n.lane_iperm = nullptr;
        const int blocks = (n.NA + 31) / 32;
        const bool pdl = h->pdl && !h->profiling;
        prof_begin(h, "k_newton_linear_lane");
        if (h->lin_lane == 32) launch_k(pdl, k_newton_linear_lane<32>, dim3(blocks), dim3(1024), h->lane_smem, cur_of(h), n, h->lane_tape, c);
        else if (h->lin_lane == 16) launch_k(pdl, k_newton_linear_lane<16>, dim3(blocks), dim3(512), h->lane_smem, cur_of(h), n, h->lane_tape, c);
        else launch_k(pdl, k_newton_linear_lane<8>, dim3(blocks), dim3(256), h->lane_smem, cur_of(h), n, h->lane_tape, c);
        prof_end(h);
        launches_of(h)++;
    } else if (h->lin_m > 0 && h->full_newton) {
        if (!getenv("RBS_TIME_SOLVE_ONLY"))
        LAUNCH(h, k_assemble_linear, (long long)(n.n_lu + n.n_reduced) * ((n.NA + 32 * RBS_MPT - 1) / (32 * RBS_MPT)) * 32, n,
               sp<int>(h, "dsrc_ptr"), sp<int>(h, "dsrc_code"));
        EntSched es = h->esched_c;
        es.smem_sched = h->lin_m_sched;
        const int M = h->lin_m, blocks = (n.NA + M - 1) / M;
        prof_begin(h, "k_newton_linear_m");
        if (M == 8) k_newton_linear_m<8><<<blocks, RBS_ENT_LANES, h->lin_m_smem, cur_of(h)>>>(n, es, c);
        else k_newton_linear_m<4><<<blocks, RBS_ENT_LANES, h->lin_m_smem, cur_of(h)>>>(n, es, c);
        prof_end(h);
        launches_of(h)++;
    } else if (h->lin_ent) {
        if (!getenv("RBS_TIME_SOLVE_ONLY"))
        LAUNCH(h, k_assemble_linear, (long long)(n.n_lu + n.n_reduced) * ((n.NA + 32 * RBS_MPT - 1) / (32 * RBS_MPT)) * 32, n,
               sp<int>(h, "dsrc_ptr"), sp<int>(h, "dsrc_code"));
        int et = h->ent_tile, blocks = (n.NA + et - 1) / et, nt = RBS_ENT_LANES * et / 2;
        int wl = h->full_newton ? 0 : 1;
        // every iteration of the full-Newton path refactors and nobody reads the factors later:
        // the compact slots serve it (more tiles per SM); the frozen path keeps all factors
        const bool cmp = h->lin_entc && h->full_newton;
        const EntSched& es = cmp ? h->esched_c : h->esched;
        const size_t sb = cmp ? h->entc_smem_bytes : h->ent_smem_bytes;
        prof_begin(h, "k_newton_linear_ent");
        const bool pdl = h->pdl && !h->profiling;
        if (et == 2) launch_k(pdl, k_newton_linear_ent<2>, dim3(blocks), dim3(nt), sb, cur_of(h), n, es, wl, c);
        else if (et == 4) launch_k(pdl, k_newton_linear_ent<4>, dim3(blocks), dim3(nt), sb, cur_of(h), n, es, wl, c);
        else launch_k(pdl, k_newton_linear_ent<8>, dim3(blocks), dim3(nt), sb, cur_of(h), n, es, wl, c);
        prof_end(h);
        launches_of(h)++;
    } else if (h->lin_smem) {
        LAUNCH(h, k_assemble_linear, (long long)(n.n_lu + n.n_reduced) * ((n.NA + 32 * RBS_MPT - 1) / (32 * RBS_MPT)) * 32, n,
               sp<int>(h, "dsrc_ptr"), sp<int>(h, "dsrc_code"));
        int blocks = (n.NA + 7) / 8;
        prof_begin(h, "k_newton_linear_smem");
        if (h->lin_smem16)
            k_newton_linear_smem<uint16_t><<<blocks, h->smem_threads, h->lin_smem_bytes, cur_of(h)>>>(
                n, h->dsched, (const uint16_t*)h->d_blob2, h->full_newton ? 0 : 1, c);
        else
            k_newton_linear_smem<int32_t><<<blocks, h->smem_threads, h->lin_smem_bytes, cur_of(h)>>>(
                n, h->dsched, (const int32_t*)h->d_blob2, h->full_newton ? 0 : 1, c);
        prof_end(h);
        launches_of(h)++;
    } else if (h->smem_tile > 0) {
        int blocks = (n.NA + h->smem_tile - 1) / h->smem_tile;
        prof_begin(h, "k_newton_linear");
#define RUN_LIN(T, IT) k_newton_linear<T, IT><<<blocks, h->smem_threads, h->smem_bytes, cur_of(h)>>>( \
            n, h->sched, (const IT*)h->d_blob, h->d_lsrc_ptr, h->d_lsrc_code, c)
        if (h->sched16) { if (h->smem_tile == 32) RUN_LIN(32, uint16_t); else if (h->smem_tile == 16) RUN_LIN(16, uint16_t); else RUN_LIN(8, uint16_t); }
        else { if (h->smem_tile == 32) RUN_LIN(32, int32_t); else if (h->smem_tile == 16) RUN_LIN(16, int32_t); else RUN_LIN(8, int32_t); }
#undef RUN_LIN
        prof_end(h);
        launches_of(h)++;
    } else {
        launch_factor<true>(h, -1.0, nw);
        launch_solve<true>(h, nullptr, c, Pred{n.mphase, PHASE_NEWTON, nullptr});
    }
}

// One pass of the per-member state machine (see rbs_kernels.cuh, "sub-stepping state machine").
// The finalize half runs on the compacted list of members in PHASE_LIMIT, the Newton half on
// the compacted list of members that are not DONE, so a pass costs what its active members cost.
static void launch_pass(rbs_handle* h, rbs_handle::Group& G, double dt, double h_min, int n_steps) {
    Net& n = net_of(h);
    cur_of(h) = G.stream;
    const int* from_ts = sp<int>(h, "ud_demand_from_timeseries");
    const double* alloc_total = sp<double>(h, "ud_alloc_total");
    // The finalize half may wait until the members whose iteration ended are at least
    // `limit_frac` of the group's unfinished members (RBS_LIMIT_FRAC, default 0 = every pass): its
    // kernels then walk a denser list (fewer launches, less sector over-fetch) at the price of idle
    // passes for the members that wait.  Same per-member arithmetic whatever the pass it runs in.
    if (G.n_limit > 0 && (double)G.n_limit >= h->limit_frac * (double)G.n_active) {
        // finalize half: members whose Newton iteration ended in the previous pass
        n.NA = G.n_limit;
        n.alist = G.d_llist;
        long long B = n.NA;
        Pred lim{n.mphase, PHASE_LIMIT, nullptr};
        LAUNCH_AS(h, "k_basin_limit_pre", (k_basin<false, 1>), (long long)n.n_reduced * B, n, nullptr, false, lim, 1);
        LAUNCH(h, k_step_limit, (long long)n.n_state * B, n, from_ts, alloc_total);
        if (h->fuse_small && G.n_limit <= 512) {  // the last block decides: two rounds of its 256 threads at most
            LAUNCH(h, k_limit_post_decide, (long long)n.n_reduced * B, n, lim, G.d_tilecnt + G.n_tiles, dt, h_min,
                   h->max_halvings, h->grow_iters, n_steps, h->shrink_log2);
        } else {
            LAUNCH_AS(h, "k_basin_limit_post", (k_basin<false, 2>), (long long)n.n_reduced * B, n, nullptr, false, lim, 2);
            LAUNCH(h, k_step_decide, B, n, dt, h_min, h->max_halvings, h->grow_iters, n_steps, h->shrink_log2);
        }
        LAUNCH(h, k_step_apply, (long long)(n.n_state + n.n_basin) * B, n, h->predictor);
        if (n.n_dc > 0) LAUNCH(h, k_discrete_control, (long long)n.n_dc * B, n, 0);
    }
    // Newton half; the last row block of a member tile in k_newton_update runs the convergence test
    n.NA = G.n_active;
    n.alist = G.d_alist;
    Pred nw{n.mphase, PHASE_NEWTON, n.mjac};
    double* c = mp(h, "cvec");
    // Small ensembles: several Newton iterations per pass.  Every kernel of the Newton half is
    // predicated on the member's phase and the convergence test runs inside k_newton_update, so
    // members whose iteration ended simply sit out the rest of the pass; the finalize half, the
    // compaction and the host round trip are then paid once per `iters_per_pass` iterations (most
    // members need three).  Same per-member arithmetic, bitwise the same results.
    for (int rep = 0; rep < h->iters_per_pass; rep++) {
        launch_rhs<true, 1>(h, nullptr, nw);
        if (!h->full_newton) launch_rhs<false, 1>(h, nullptr, nw);
        launch_linear(h, c, nw);
        dim3 block(32, RBS_DZ_ROWS), grid((n.NA + 63) / 64, h->n_part);
        ConvArgs cv{G.d_tilecnt, h->n_part, h->max_iter, h->full_newton, h->early_exit, 0.01};
        LAUNCH_CFG(h, k_newton_update, grid, block, n, c, h->abstol, h->reltol, h->rows_per_block, cv);
    }
    n.NA = n.B;
    n.alist = nullptr;
    LAUNCH_CFG(h, k_compact2, 1, 1024, n, G.lo, G.hi, G.d_alist, G.d_llist, G.d_counts);
    if (!tl_ctx) h->cur = h->stream;
    passes_of(h)++;
}

}  // extern "C"
template <int CL>
static cudaError_t launch_finalize_cl(cudaStream_t st, int tiles, const Net& n, const FinArgs& a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tiles * CL));
    cfg.blockDim = dim3(RBS_FIN_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_finalize<CL>, n, a);
}
extern "C" {

// Round-2 pass of a member group: one Newton iteration of every member that is not done, then the
// convergence test and — for the members whose iteration ended — the whole finalize half in one
// cluster kernel, then the work list of the next pass.
static int launch_pass_r2(rbs_handle* h, rbs_handle::Group& G, double dt, double h_min, int n_steps) {
    Net& n = h->net;
    h->cur = G.stream;
    n.NA = G.n_active;
    n.alist = G.d_alist;
    Pred nw{n.mphase, PHASE_NEWTON, n.mjac};
    launch_rhs<true, 1>(h, nullptr, nw);
    if (!h->full_newton) launch_rhs<false, 1>(h, nullptr, nw);
    double* c = mp(h, "cvec");
    launch_linear(h, c, nw);
    {
        dim3 block(32, RBS_DZ_ROWS), grid((n.NA + 63) / 64, h->n_part);
        LAUNCH_CFG(h, k_newton_update, grid, block, n, c, h->abstol, h->reltol, h->rows_per_block, ConvArgs{});
    }
    n.NA = n.B;
    n.alist = nullptr;
    FinArgs a{};
    a.from_ts = sp<int>(h, "ud_demand_from_timeseries");
    a.alloc_total = sp<double>(h, "ud_alloc_total");
    a.dt = dt; a.h_min = h_min; a.kappa = 0.01;
    a.max_halvings = h->max_halvings; a.grow_iters = h->grow_iters; a.n_steps = n_steps;
    a.shrink_log2 = h->shrink_log2; a.predictor = h->predictor;
    a.n_part = h->n_part; a.max_iter = h->max_iter; a.full_newton = h->full_newton; a.early_exit = h->early_exit;
    a.m_lo = G.lo; a.m_hi = G.hi;
    const int tiles = (G.hi - G.lo + 31) / 32;
    prof_begin(h, "k_finalize");
    cudaError_t e;
    switch (h->fin_cluster) {
        case 8: e = launch_finalize_cl<8>(G.stream, tiles, n, a); break;
        case 4: e = launch_finalize_cl<4>(G.stream, tiles, n, a); break;
        case 2: e = launch_finalize_cl<2>(G.stream, tiles, n, a); break;
        default: e = launch_finalize_cl<1>(G.stream, tiles, n, a); break;
    }
    prof_end(h);
    h->launches++;
    if (e != cudaSuccess) { h->cur = h->stream; return fail(std::string("k_finalize launch: ") + cudaGetErrorString(e)); }
    LAUNCH_CFG(h, k_compact_tiles, 1, 1024, n, G.lo, G.hi, G.d_alist, G.d_counts);
    h->cur = h->stream;
    h->n_passes++;
    return 0;
}

int rbs_time_linear(rbs_handle* h, int32_t n_active, int32_t reps, double* ms_per_launch) {
    if (!h || !ms_per_launch) return fail("rbs_time_linear: null argument");
    if (n_active < 1 || n_active > h->B || reps < 1) return fail("rbs_time_linear: invalid argument");
    cudaSetDevice(h->device);
    Net& n = h->net;
    // all members in the Newton phase with a Jacobian refresh, sub-step of 1 hour
    std::vector<int> ph(n.B, PHASE_NEWTON), one(n.B, 1);
    std::vector<double> hh(n.B, 3600.0);
    CUDA_TRY(cudaMemcpyAsync(n.mphase, ph.data(), n.B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(n.mjac, one.data(), n.B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(n.mh, hh.data(), n.B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    n.NA = n_active;
    n.alist = nullptr;
    Pred nw{n.mphase, PHASE_NEWTON, n.mjac};
    double* c = mp(h, "cvec");
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch_linear(h, c, nw);  // warm-up
    cudaEventRecord(a, h->stream);
    for (int r = 0; r < reps; r++) launch_linear(h, c, nw);
    cudaEventRecord(b, h->stream);
    n.NA = n.B;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_per_launch = ms / reps;
#ifdef RBS_LIN_PROFILE
    if (h->lin_ent) {
        long long lv[64], clk[8];
        cudaMemcpyFromSymbol(lv, g_lin_lvl, sizeof lv);
        cudaMemcpyFromSymbol(clk, g_lin_clk, sizeof clk);
        int nl = h->esched.n_elim + h->esched.n_bwd;
        fprintf(stderr, "levels (cycles):");
        for (int l = 0; l < nl && l < 64; l++) fprintf(stderr, " %lld", lv[l] - (l ? lv[l - 1] : clk[2]));
        fprintf(stderr, "\n");
    }
    if (h->lin_smem) {
        long long clk[8];
        cudaMemcpyFromSymbol(clk, g_lin_clk, sizeof clk);
        fprintf(stderr, "lin phases (cycles): blob %lld init %lld elim %lld rhs %lld fwd %lld bwd %lld out %lld total %lld\n",
                clk[1] - clk[0], clk[2] - clk[1], clk[3] - clk[2], clk[4] - clk[3], clk[5] - clk[4],
                clk[6] - clk[5], clk[7] - clk[6], clk[7] - clk[0]);
    }
#endif
    return check_async(h, "rbs_time_linear");
}

// the window as one cooperative kernel (small ensembles, full Newton)
static int advance_coop(rbs_handle* h, double dt, double h_min, int32_t n_steps) {
    Net& n = h->net;
    n.NA = n.B;
    n.alist = nullptr;
    CoopArgs a{};
    a.from_ts = sp<int>(h, "ud_demand_from_timeseries");
    a.alloc_total = sp<double>(h, "ud_alloc_total");
    a.dsrc_ptr = sp<int>(h, "dsrc_ptr");
    a.dsrc_code = sp<int>(h, "dsrc_code");
    a.list_cc = h->d_list_phase[1]; a.n_list_cc = (int)h->list_phase[1].size();
    a.list_pid = h->d_list_phase[2]; a.n_list_pid = (int)h->list_phase[2].size();
    a.pid_fused = h->pid_fused ? 1 : 0;
    a.c = mp(h, "cvec");
    a.dt = dt; a.h_min = h_min; a.abstol = h->abstol; a.reltol = h->reltol;
    a.max_halvings = h->max_halvings; a.grow_iters = h->grow_iters; a.n_steps = n_steps;
    a.shrink_log2 = h->shrink_log2; a.predictor = h->predictor; a.early_exit = h->early_exit;
    a.max_iter = h->max_iter; a.n_part = h->n_part; a.rows_per_block = h->rows_per_block;
    a.max_passes = std::min<long long>(h->max_passes, 4000) * (long long)n_steps;  // safety net
    a.counters = h->d_coop_counters;
    CUDA_TRY(cudaMemsetAsync(h->d_coop_counters, 0, 4 * sizeof(int), h->stream));
    EntSched es = h->lin_entc ? h->esched_c : h->esched;
    es.smem_sched = 0;
    size_t sb2 = (size_t)es.n_slots * 2 * sizeof(double);
    // no more blocks than there is work in the widest phase
    long long widest = (long long)std::max(n.n_state + n.n_basin, n.n_lu + n.n_reduced) * n.B;
    int blocks = (int)std::max<long long>(1, std::min<long long>(h->coop_blocks, (widest + RBS_COOP_THREADS - 1) / RBS_COOP_THREADS));
    void* args[] = {&n, &es, &a};
    prof_begin(h, "k_window_coop");
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_window_coop, dim3(blocks), dim3(RBS_COOP_THREADS), args, sb2, h->stream));
    prof_end(h);
    h->launches++;
    CUDA_TRY(cudaMemcpyAsync(h->h_coop_counters, h->d_coop_counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const int np_ = h->h_coop_counters[3];
    if (np_ >= a.max_passes && h->h_coop_counters[(np_ + 2) % 3] != 0)
        return fail("rbs_advance: pass limit exceeded");
    h->n_passes += np_;
    return 0;
}

// the window as one kernel with a block per member tile (full Newton)
static int advance_tile(rbs_handle* h, double dt, double h_min, int32_t n_steps) {
    Net& n = h->net;
    n.NA = n.B;
    n.alist = nullptr;
    CoopArgs a{};
    a.from_ts = sp<int>(h, "ud_demand_from_timeseries");
    a.alloc_total = sp<double>(h, "ud_alloc_total");
    a.dsrc_ptr = sp<int>(h, "dsrc_ptr");
    a.dsrc_code = sp<int>(h, "dsrc_code");
    a.list_cc = h->d_list_phase[1]; a.n_list_cc = (int)h->list_phase[1].size();
    a.list_pid = h->d_list_phase[2]; a.n_list_pid = (int)h->list_phase[2].size();
    a.pid_fused = h->pid_fused ? 1 : 0;
    a.c = mp(h, "cvec");
    a.dt = dt; a.h_min = h_min; a.abstol = h->abstol; a.reltol = h->reltol;
    a.max_halvings = h->max_halvings; a.grow_iters = h->grow_iters; a.n_steps = n_steps;
    a.shrink_log2 = h->shrink_log2; a.predictor = h->predictor; a.early_exit = h->early_exit;
    a.max_iter = h->max_iter; a.n_part = h->n_part; a.rows_per_block = h->rows_per_block;
    a.max_passes = std::min<long long>(h->max_passes, 4000) * (long long)n_steps;  // safety net
    a.counters = h->d_coop_counters;
    CUDA_TRY(cudaMemsetAsync(h->d_coop_counters, 0, 4 * sizeof(int), h->stream));
    EntSched es = h->lin_entc ? h->esched_c : h->esched;
    es.smem_sched = 0;
    const int tw = h->tile_window, blocks = (n.B + tw - 1) / tw, threads = RBS_ENT_LANES * tw / 2;
    const size_t sbt = std::max<size_t>((size_t)es.n_slots, (size_t)h->n_part * RBS_DZ_ROWS) * tw * sizeof(double);
    prof_begin(h, "k_window_tile");
    if (tw == 2) k_window_tile<2><<<blocks, threads, sbt, h->stream>>>(n, es, a);
    else k_window_tile<4><<<blocks, threads, sbt, h->stream>>>(n, es, a);
    prof_end(h);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    CUDA_TRY(cudaMemcpyAsync(h->h_coop_counters, h->d_coop_counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->h_coop_counters[0] != 0) return fail("rbs_advance: pass limit exceeded");
    h->n_passes += h->h_coop_counters[3];
    return 0;
}

// result slices of the outer steps that every member of every group has completed go out while
// the stragglers are still computing (every kernel that wrote them has been observed complete)
static int stream_finished_slices(rbs_handle* h, int n_steps) {
    if (!h->wout.on) return 0;
    int ready = n_steps;
    for (auto& G2 : h->groups) ready = std::min(ready, G2.min_step);
    if (ready > h->wout.streamed) {
        auto& w = h->wout;
        int k0 = w.streamed, cnt = ready - k0;
        if (w.dev_s)
            CUDA_TRY(cudaMemcpyAsync((char*)w.host_s + w.slice_s * k0, (char*)w.dev_s + w.slice_s * k0,
                                     w.slice_s * cnt, cudaMemcpyDeviceToHost, h->d2h_stream));
        if (w.dev_f)
            CUDA_TRY(cudaMemcpyAsync((char*)w.host_f + w.slice_f * k0, (char*)w.dev_f + w.slice_f * k0,
                                     w.slice_f * cnt, cudaMemcpyDeviceToHost, h->d2h_stream));
        w.streamed = ready;
    }
    return 0;
}

// Pass loop of the stream path: a pass per member group and round, the groups on their own streams;
// the host reads a group's list lengths (12 bytes) after each of its passes to size the next one.
// (Measured in round 2, profiles/summary_r2.md: list lengths read by the kernels from device
// memory, so that passes are launched without waiting, cost the latency-bound kernels 5-15 %
// each — one more dependent load at the head of every thread's chain, or surplus blocks — and win
// nothing, because the host round trip of one group hides behind the kernels of the others.)
static int advance_passes(rbs_handle* h, double dt, double h_min, int32_t n_steps) {
    Net& n = h->net;
    // safety net only: every attempt either advances the member or halves its sub-step
    const int64_t max_passes = h->max_passes * n_steps * (int64_t)h->groups.size();
    int64_t passes = 0;
    size_t n_done = 0;
    static const bool trace = getenv("RBS_TRACE_PASSES") != nullptr;
    for (auto& G : h->groups) {
        if (G.stream != h->stream) CUDA_TRY(cudaStreamWaitEvent(G.stream, h->groups[0].ev, 0));
        G.n_active = G.hi - G.lo;
        G.n_limit = 0;
        G.pending = false;
        G.done = G.n_active == 0;
        G.min_step = G.done ? n_steps : 0;
        n_done += G.done;
        if (G.done) continue;
        h->cur = G.stream;
        CUDA_TRY(cudaMemsetAsync(G.d_counts, 0, 4 * sizeof(int), G.stream));
        LAUNCH_CFG(h, k_compact2, 1, 1024, n, G.lo, G.hi, G.d_alist, G.d_llist, G.d_counts);
    }
    h->cur = h->stream;
    while (n_done < h->groups.size()) {
        bool progressed = false;
        for (auto& G : h->groups) {
            if (G.done) continue;
            if (G.pending) {
                cudaError_t q = cudaEventQuery(G.ev);
                if (q == cudaErrorNotReady) continue;
                if (q != cudaSuccess) return fail(std::string("rbs_advance: ") + cudaGetErrorString(q));
                G.pending = false;
                G.n_active = G.h_counts[0];
                G.n_limit = G.h_counts[1];
                G.min_step = G.h_counts[2];
                if (trace) fprintf(stderr, "pass g%d %lld active %d limit %d min_step %d\n", (int)(&G - &h->groups[0]),
                                   (long long)passes, G.n_active, G.n_limit, G.min_step);
                if (stream_finished_slices(h, n_steps)) return 1;
                if (G.n_active == 0) { G.done = true; n_done++; progressed = true; continue; }
            }
            if (++passes > max_passes) return fail("rbs_advance: pass limit exceeded");
            launch_pass(h, G, dt, h_min, n_steps);
            CUDA_TRY(cudaMemcpyAsync(G.h_counts, G.d_counts, 3 * sizeof(int), cudaMemcpyDeviceToHost, G.stream));
            CUDA_TRY(cudaEventRecord(G.ev, G.stream));
            G.pending = true;
            progressed = true;
        }
        if (!progressed && h->groups.size() == 1) cudaEventSynchronize(h->groups[0].ev);
    }
    return 0;
}

// The same pass loop with one host thread per member group: a group's next pass is launched the
// moment its counts arrive, instead of when the single host thread gets back to it after launching
// the other groups' passes (a dozen launches each).  Every thread launches through its own
// LaunchCtx (kernel argument copy + stream); shared bookkeeping is under a mutex.
static int group_worker(rbs_handle* h, int gi, double dt, double h_min, int32_t n_steps, int64_t max_passes,
                        std::mutex& mu, int& rc, std::string& err) {
    cudaSetDevice(h->device);
    rbs_handle::Group& G = h->groups[gi];
    LaunchCtx ctx;
    ctx.net = h->net;
    ctx.cur = G.stream;
    tl_ctx = &ctx;
    static const bool trace = getenv("RBS_TRACE_PASSES") != nullptr;
    int my_rc = 0;
    auto body = [&]() -> int {
        int64_t passes = 0;
        while (!G.done) {
            { std::lock_guard<std::mutex> lk(mu); if (rc) return 0; }
            if (++passes > max_passes) return fail("rbs_advance: pass limit exceeded");
            launch_pass(h, G, dt, h_min, n_steps);
            CUDA_TRY(cudaMemcpyAsync(G.h_counts, G.d_counts, 3 * sizeof(int), cudaMemcpyDeviceToHost, G.stream));
            CUDA_TRY(cudaEventRecord(G.ev, G.stream));
            CUDA_TRY(cudaEventSynchronize(G.ev));
            G.n_active = G.h_counts[0];
            G.n_limit = G.h_counts[1];
            if (trace) fprintf(stderr, "pass g%d %lld active %d limit %d min_step %d\n", gi, (long long)passes,
                               G.n_active, G.n_limit, G.h_counts[2]);
            {
                std::lock_guard<std::mutex> lk(mu);
                G.min_step = G.h_counts[2];
                if (stream_finished_slices(h, n_steps)) return 1;
            }
            if (G.n_active == 0) G.done = true;
        }
        return 0;
    };
    my_rc = body();
    {
        std::lock_guard<std::mutex> lk(mu);
        h->launches += ctx.launches;
        h->n_passes += ctx.n_passes;
        if (my_rc && !rc) { rc = my_rc; err = g_err; }
    }
    tl_ctx = nullptr;
    return my_rc;
}

static int advance_passes_mt(rbs_handle* h, double dt, double h_min, int32_t n_steps) {
    Net& n = h->net;
    const int64_t max_passes = h->max_passes * n_steps;
    for (auto& G : h->groups) {
        if (G.stream != h->stream) CUDA_TRY(cudaStreamWaitEvent(G.stream, h->groups[0].ev, 0));
        G.n_active = G.hi - G.lo;
        G.n_limit = 0;
        G.pending = false;
        G.done = G.n_active == 0;
        G.min_step = G.done ? n_steps : 0;
        if (G.done) continue;
        h->cur = G.stream;
        CUDA_TRY(cudaMemsetAsync(G.d_counts, 0, 4 * sizeof(int), G.stream));
        LAUNCH_CFG(h, k_compact2, 1, 1024, n, G.lo, G.hi, G.d_alist, G.d_llist, G.d_counts);
    }
    h->cur = h->stream;
    std::mutex mu;
    int rc = 0;
    std::string err;
    std::vector<std::thread> workers;
    for (int gi = 1; gi < (int)h->groups.size(); gi++)
        workers.emplace_back(group_worker, h, gi, dt, h_min, n_steps, max_passes, std::ref(mu), std::ref(rc), std::ref(err));
    group_worker(h, 0, dt, h_min, n_steps, max_passes, mu, rc, err);
    for (auto& w : workers) w.join();
    if (rc) return fail(err);
    return 0;
}

// round-2 experiment kept for comparison (RBS_R2_PASS=1): cluster finalize kernel, one host
// round trip per pass
static int advance_sync_r2(rbs_handle* h, double dt, double h_min, int32_t n_steps) {
    Net& n = h->net;
    for (auto& G : h->groups) {
        if (G.stream != h->stream) CUDA_TRY(cudaStreamWaitEvent(G.stream, h->groups[0].ev, 0));
        h->cur = G.stream;
        LAUNCH_CFG(h, k_compact_tiles, 1, 1024, n, G.lo, G.hi, G.d_alist, G.d_counts);
        G.n_active = G.hi - G.lo;
        G.n_limit = 0;
        G.pending = false;
        G.done = G.n_active == 0;
        G.min_step = G.done ? n_steps : 0;
    }
    h->cur = h->stream;
    const int64_t max_passes = h->max_passes * n_steps * (int64_t)h->groups.size();
    int64_t passes = 0;
    size_t n_done = 0;
    for (auto& G : h->groups) n_done += G.done;
    while (n_done < h->groups.size()) {
        bool progressed = false;
        for (auto& G : h->groups) {
            if (G.done) continue;
            if (G.pending) {
                cudaError_t q = cudaEventQuery(G.ev);
                if (q == cudaErrorNotReady) continue;
                if (q != cudaSuccess) return fail(std::string("rbs_advance: ") + cudaGetErrorString(q));
                G.pending = false;
                G.n_active = G.h_counts[0];
                G.n_limit = G.h_counts[1];
                G.min_step = G.h_counts[2];
                if (stream_finished_slices(h, n_steps)) return 1;
                if (G.n_active == 0) { G.done = true; n_done++; progressed = true; continue; }
            }
            if (++passes > max_passes) return fail("rbs_advance: pass limit exceeded");
            if (launch_pass_r2(h, G, dt, h_min, n_steps)) return 1;
            CUDA_TRY(cudaMemcpyAsync(G.h_counts, G.d_counts, 3 * sizeof(int), cudaMemcpyDeviceToHost, G.stream));
            CUDA_TRY(cudaEventRecord(G.ev, G.stream));
            G.pending = true;
            progressed = true;
        }
        if (!progressed && h->groups.size() == 1) cudaEventSynchronize(h->groups[0].ev);
    }
    return 0;
}

static int advance_impl(rbs_handle* h, double t, double dt, int32_t n_steps, int32_t* status_per_member) {
    cudaSetDevice(h->device);
    Net& n = h->net;
    long long B = n.B;
    double h_min = std::ldexp(dt, -h->max_halvings);
    h->cur = h->stream;
    n.win_t0 = t; n.win_dt = dt;
    LAUNCH(h, k_step_begin, (long long)(n.n_state + n.n_basin + 1) * B, n, dt, h->predictor);
    const bool use_coop = h->coop_blocks > 0 && n.B <= h->coop_members;
    if ((use_coop || h->tile_window > 0) && h->full_newton && !h->profiling) {
        if (use_coop ? advance_coop(h, dt, h_min, n_steps) : advance_tile(h, dt, h_min, n_steps)) return 1;
        h->tprev = t + dt * n_steps;
        h->n_steps += n_steps;
        if (commit_staged(h)) return 1;
        if (status_per_member) {
            CUDA_TRY(cudaMemcpyAsync(status_per_member, n.status, (size_t)n.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            for (int m = 0; m < n.B; m++) if (status_per_member[m] & 1) h->n_failed++;
        }
        return check_async(h, "rbs_advance");
    }
    // fork: every group starts from the common begin on its own stream
    cudaEvent_t ev_begin = h->groups[0].ev;
    CUDA_TRY(cudaEventRecord(ev_begin, h->stream));
    if (h->r2_pass) { if (advance_sync_r2(h, dt, h_min, n_steps)) return 1; }
    else if (h->host_threads && h->groups.size() > 1 && !h->profiling) { if (advance_passes_mt(h, dt, h_min, n_steps)) return 1; }
    else if (advance_passes(h, dt, h_min, n_steps)) return 1;
    // join: every group's last event has been observed complete by the host, so all group
    // streams are idle and later work enqueued on the handle's stream is ordered after them
    h->tprev = t + dt * n_steps;
    h->n_steps += n_steps;
    // uploads staged while this window was running apply from the next window on
    if (commit_staged(h)) return 1;
    if (status_per_member) {
        CUDA_TRY(cudaMemcpyAsync(status_per_member, n.status, (size_t)n.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int m = 0; m < n.B; m++) if (status_per_member[m] & 1) h->n_failed++;
    }
    return check_async(h, "rbs_advance");
}

int rbs_advance(rbs_handle* h, double t, double dt, int32_t n_steps, int32_t* status_per_member) {
    if (!h) return fail("rbs_advance: null handle");
    if (!(dt > 0) || n_steps < 1) return fail("rbs_advance: dt must be positive and n_steps >= 1");
    return advance_impl(h, t, dt, n_steps, status_per_member);
}

static const char* const FW_KEYS[6] = {"precipitation", "surface_runoff", "potential_evaporation",
                                       "drainage", "infiltration", "fb_rate"};

int rbs_stage_forcing_window(rbs_handle* h, const char* key, const double* pinned_host,
                             int64_t n_entity, int32_t first_step, int32_t n_steps) {
    if (!h || !key || !pinned_host) return fail("rbs_stage_forcing_window: null argument");
    if (n_steps < 1 || first_step < 0) return fail("rbs_stage_forcing_window: invalid step range");
    cudaSetDevice(h->device);
    int k = -1;
    for (int i = 0; i < 6; i++) if (!strcmp(key, FW_KEYS[i])) k = i;
    if (k < 0) return fail(std::string("rbs_stage_forcing_window: not a forcing array: ") + key);
    if (h->member_entities[key] != n_entity) return fail(std::string("member array size mismatch: ") + key);
    if (n_entity == 0) return 0;
    size_t slice = (size_t)n_entity * h->B * sizeof(double);
    size_t need = slice * (size_t)(first_step + n_steps);
    if (need > h->fw_bytes[k]) {
        // grow both buffers of this array, keeping what they hold (slices staged so far, slices
        // committed for the next window)
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
        size_t cap = std::max(need, slice * 8);
        for (double** buf : {&h->fw_live[k], &h->fw_shadow[k]}) {
            double* nb_ = nullptr;
            CUDA_TRY(cudaMalloc(&nb_, cap));
            if (*buf) {
                CUDA_TRY(cudaMemcpy(nb_, *buf, h->fw_bytes[k], cudaMemcpyDeviceToDevice));
                cudaFree(*buf);
            }
            *buf = nb_;
        }
        h->fw_bytes[k] = cap;
    }
    // the previous window's tail (copy of its last slice into the current arrays) reads the buffer
    // that is the shadow now
    CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_win, 0));
    CUDA_TRY(cudaMemcpyAsync((char*)h->fw_shadow[k] + slice * first_step, pinned_host, slice * n_steps,
                             cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(cudaEventRecord(h->ev_h2d, h->copy_stream));
    h->fw_staged_steps[k] = std::max(h->fw_staged_steps[k], first_step + n_steps);
    return 0;
}

// staged slices -> live (the shadow buffers take the next upload)
static int commit_window(rbs_handle* h) {
    bool any = false;
    for (int k = 0; k < 6; k++) {
        if (h->fw_staged_steps[k] == 0) continue;
        std::swap(h->fw_live[k], h->fw_shadow[k]);
        h->fw_live_steps[k] = h->fw_staged_steps[k];
        h->fw_staged_steps[k] = 0;
        any = true;
    }
    if (any) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_h2d, 0));
    return 0;
}

int rbs_set_forcing_slicing(rbs_handle* h, int32_t steps_per_slice, int32_t phase) {
    if (!h) return fail("rbs_set_forcing_slicing: null handle");
    if (steps_per_slice < 1 || phase < 0 || phase >= steps_per_slice)
        return fail("rbs_set_forcing_slicing: need steps_per_slice >= 1 and 0 <= phase < steps_per_slice");
    h->fw_every = steps_per_slice;
    h->fw_phase = phase;
    return 0;
}

int rbs_commit_forcing_window(rbs_handle* h) {
    if (!h) return fail("rbs_commit_forcing_window: null handle");
    cudaSetDevice(h->device);
    return commit_window(h);
}

int rbs_advance_window(rbs_handle* h, double t, double dt, int32_t n_steps,
                       double* storage_out_pinned, double* flow_out_pinned,
                       int32_t* status_per_member) {
    if (!h) return fail("rbs_advance_window: null handle");
    if (!(dt > 0) || n_steps < 1) return fail("rbs_advance_window: dt must be positive and n_steps >= 1");
    cudaSetDevice(h->device);
    Net& n = h->net;
    const double** wp[6] = {&n.w_precipitation, &n.w_surface_runoff, &n.w_potential_evaporation,
                            &n.w_drainage, &n.w_infiltration, &n.w_fb_rate};
    n.fw_every = h->fw_every; n.fw_phase = h->fw_phase;
    const int last_slice = (h->fw_phase + n_steps - 1) / h->fw_every, n_slices = last_slice + 1;
    for (int k = 0; k < 6; k++) {
        *wp[k] = nullptr;
        if (h->fw_live_steps[k] == 0) continue;
        if (h->fw_live_steps[k] < n_slices)
            return fail(std::string("rbs_advance_window: fewer steps staged than advanced: ") + FW_KEYS[k]);
        *wp[k] = h->fw_live[k];
    }
    n.fw_sb = (long long)n.n_basin * n.B;
    n.fw_sf = (long long)n.n_fb * n.B;
    n.w_storage = nullptr;
    int cur = h->fw_res_cur;
    size_t res_bytes = (size_t)n_steps * n.n_basin * n.B * sizeof(double);
    if (storage_out_pinned && n.n_basin > 0) {
        if (res_bytes > h->fw_res_bytes) {
            CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
            for (int k = 0; k < 2; k++) {
                if (h->fw_res[k]) cudaFree(h->fw_res[k]);
                h->fw_res[k] = nullptr;
                CUDA_TRY(cudaMalloc(&h->fw_res[k], res_bytes));
            }
            h->fw_res_bytes = res_bytes;
        }
        n.w_storage = h->fw_res[cur];
    }
    n.w_flow = nullptr;
    n.fw_dt = dt;
    size_t flow_bytes = (size_t)n_steps * n.n_state * n.B * sizeof(double);
    if (flow_out_pinned) {
        if (flow_bytes > h->fw_flow_bytes) {
            CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
            for (int k = 0; k < 2; k++) {
                if (h->fw_flow[k]) cudaFree(h->fw_flow[k]);
                h->fw_flow[k] = nullptr;
                CUDA_TRY(cudaMalloc(&h->fw_flow[k], flow_bytes));
            }
            h->fw_flow_bytes = flow_bytes;
        }
        n.w_flow = h->fw_flow[cur];
        n.ustep = mp(h, "ustep");
    }
    // the downloads of the window before last used these buffers
    if (n.w_storage || n.w_flow) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_d2h[cur], 0));
    n.fw_on = 1;
    h->wout.on = n.w_storage || n.w_flow;
    h->wout.host_s = storage_out_pinned; h->wout.dev_s = n.w_storage;
    h->wout.host_f = flow_out_pinned; h->wout.dev_f = n.w_flow;
    h->wout.slice_s = (size_t)n.n_basin * n.B * sizeof(double);
    h->wout.slice_f = (size_t)n.n_state * n.B * sizeof(double);
    h->wout.streamed = 0;
    int rc = advance_impl(h, t, dt, n_steps, status_per_member);
    n.fw_on = 0;
    h->wout.on = false;
    if (rc) return rc;
    // the last slice stays in force (and readable through rbs_get_member_f64) after the window
    for (int k = 0; k < 6; k++) {
        if (!*wp[k]) continue;
        int64_t ne = h->member_entities[FW_KEYS[k]];
        size_t slice = (size_t)ne * n.B * sizeof(double);
        CUDA_TRY(cudaMemcpyAsync(h->member[FW_KEYS[k]].p, (const char*)h->fw_live[k] + slice * last_slice, slice,
                                 cudaMemcpyDeviceToDevice, h->stream));
        *wp[k] = nullptr;
        h->fw_live_steps[k] = 0;  // a window of slices serves one rbs_advance_window
    }
    CUDA_TRY(cudaEventRecord(h->ev_win, h->stream));
    if (n.w_storage || n.w_flow) {
        // what the progressive download has not taken yet
        CUDA_TRY(cudaStreamWaitEvent(h->d2h_stream, h->ev_win, 0));
        const int k0 = h->wout.streamed, cnt = n_steps - k0;
        if (n.w_storage && cnt > 0)
            CUDA_TRY(cudaMemcpyAsync((char*)storage_out_pinned + h->wout.slice_s * k0, (char*)h->fw_res[cur] + h->wout.slice_s * k0,
                                     h->wout.slice_s * cnt, cudaMemcpyDeviceToHost, h->d2h_stream));
        if (n.w_flow && cnt > 0)
            CUDA_TRY(cudaMemcpyAsync((char*)flow_out_pinned + h->wout.slice_f * k0, (char*)h->fw_flow[cur] + h->wout.slice_f * k0,
                                     h->wout.slice_f * cnt, cudaMemcpyDeviceToHost, h->d2h_stream));
        CUDA_TRY(cudaEventRecord(h->ev_d2h[cur], h->d2h_stream));
        h->fw_res_cur = cur ^ 1;
        n.w_storage = nullptr;
        n.w_flow = nullptr;
    }
    // slices staged while this window was running serve the next window
    if (commit_window(h)) return 1;
    return check_async(h, "rbs_advance_window");
}

int rbs_step_implicit_euler(rbs_handle* h, double t, double dt, int32_t* status_per_member) {
    return rbs_advance(h, t, dt, 1, status_per_member);
}

}  // extern "C"
