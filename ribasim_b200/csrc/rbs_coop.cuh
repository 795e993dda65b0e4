// rbs_coop.cuh — the whole window as ONE cooperative kernel (small ensembles, single instances).
//
// With a few hundred members a pass of the stream path is a chain of a dozen kernels of a few
// microseconds each plus a host round trip for the work lists: the GPU idles between launches.
// Here one persistent grid (as many blocks as are co-resident) runs the passes itself: every
// phase is a grid-stride loop over (entity, member) items calling the same device functions as
// the kernels of the stream path, phases are separated by grid barriers, the member state
// machine (rbs_kernels.cuh, "sub-stepping state machine") is evaluated with predicates instead
// of compacted work lists, and the loop ends when no member is left (integer counter).
// Per-member arithmetic, summation orders and the accept / reject logic are those of the stream
// path: results are bitwise identical.  Full-Newton mode only (the stream path also serves the
// frozen variant).
#pragma once
#include <cooperative_groups.h>

#include "rbs_kernels.cuh"

struct CoopArgs {
    const int* from_ts;
    const double* alloc_total;
    const int *dsrc_ptr, *dsrc_code;
    const int *list_cc, *list_pid;
    int n_list_cc, n_list_pid, pid_fused;
    double* c;
    double dt, h_min, abstol, reltol;
    int max_halvings, grow_iters, n_steps, shrink_log2, predictor, early_exit, max_iter, n_part,
        rows_per_block;
    long long max_passes;
    int* counters;  // [0..2]: members left after a pass (rotating); [3]: passes executed
};

#define RBS_COOP_THREADS 256
// (entity, member) items of a phase, member fastest: consecutive threads = consecutive members
#define RBS_COOP_FOR(count)                                                                   \
    for (long long it_ = gtid; it_ < (long long)(count) * B; it_ += gsize)                    \
        for (int ent = (int)(it_ / B), m = (int)(it_ - (long long)ent * B), once_ = 1; once_; once_ = 0)

__global__ void __launch_bounds__(RBS_COOP_THREADS, 2) k_window_coop(Net n, EntSched es, CoopArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem[];
    __shared__ double sh_upd[2][RBS_DZ_ROWS][33];
    const long long gsize = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int B = n.B;
    int passes = 0;
    for (long long pass = 0; pass < a.max_passes; pass++) {
        // three rotating counters: the one zeroed here was last read two passes ago, with grid
        // barriers in between (zeroing the previous pass's counter would race with its readers)
        int* left = a.counters + (int)(pass % 3);
        if (gtid == 0) a.counters[(int)((pass + 1) % 3)] = 0;
        // ---- finalize half: members whose Newton iteration ended in the previous pass
        RBS_COOP_FOR(n.n_reduced)
            if (n.mphase[m] == PHASE_LIMIT) rbs_basin_item<false, 1>(n, nullptr, false, 1, ent, m);
        grid.sync();
        RBS_COOP_FOR(n.n_state) rbs_step_limit_item(n, a.from_ts, a.alloc_total, ent, m);
        grid.sync();
        RBS_COOP_FOR(n.n_reduced)
            if (n.mphase[m] == PHASE_LIMIT) rbs_basin_item<false, 2>(n, nullptr, false, 2, ent, m);
        grid.sync();
        for (long long m = gtid; m < B; m += gsize)
            rbs_step_decide_item(n, a.dt, a.h_min, a.max_halvings, a.grow_iters, a.n_steps, a.shrink_log2, (int)m);
        grid.sync();
        RBS_COOP_FOR(n.n_state + n.n_basin) rbs_step_apply_item(n, a.predictor, ent, m);
        if (n.n_dc > 0) {
            grid.sync();  // the control state is read by the links of the Newton half
            RBS_COOP_FOR(n.n_dc) rbs_discrete_control_item(n, 0, ent, m);
        }
        grid.sync();
        // ---- Newton half: members in PHASE_NEWTON, Jacobian refreshed every iteration
        RBS_COOP_FOR(n.n_reduced)
            if (n.mphase[m] == PHASE_NEWTON) rbs_basin_item<true, 1>(n, nullptr, true, 0, ent, m);
        grid.sync();
        RBS_COOP_FOR(n.n_conn)
            if (n.mphase[m] == PHASE_NEWTON) rbs_link_item<true>(n, ent, m, 0);
        grid.sync();
        if (n.n_cc > 0) {
            RBS_COOP_FOR(n.n_cc)
                if (n.mphase[m] == PHASE_NEWTON) rbs_cc_item<true>(n, ent, m);
            grid.sync();
            RBS_COOP_FOR(a.n_list_cc)
                if (n.mphase[m] == PHASE_NEWTON) rbs_link_item<true>(n, a.list_cc[ent], m, CONTROL_CONTINUOUS);
            grid.sync();
        }
        if (n.n_pid > 0) {
            RBS_COOP_FOR(n.n_pid)
                if (n.mphase[m] == PHASE_NEWTON) rbs_pid_item<true>(n, n.ur, a.pid_fused, ent, m);
            grid.sync();
            if (!a.pid_fused) {
                RBS_COOP_FOR(a.n_list_pid)
                    if (n.mphase[m] == PHASE_NEWTON) rbs_link_item<true>(n, a.list_pid[ent], m, CONTROL_PID);
                grid.sync();
            }
        }
        RBS_COOP_FOR(n.n_lu + n.n_reduced)
            if (n.mphase[m] == PHASE_NEWTON) rbs_assemble_item(n, a.dsrc_ptr, a.dsrc_code, ent, m);
        grid.sync();
        for (int tile = blockIdx.x; tile * 2 < B; tile += gridDim.x) {  // block-uniform
            rbs_lin_tile<2>(n, es, smem, tile * 2, 0, a.c);
            __syncthreads();  // the slots are reused by the block's next tile
        }
        grid.sync();
        {
            const int n_t64 = (B + 63) >> 6, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
            for (int unit = blockIdx.x; unit < n_t64 * a.n_part; unit += gridDim.x) {  // block-uniform
                rbs_update_unit(n, a.c, a.abstol, a.reltol, a.rows_per_block, unit % n_t64, unit / n_t64, tx, ty, sh_upd);
                __syncthreads();
            }
        }
        grid.sync();
        for (long long m = gtid; m < B; m += gsize) {
            if (n.mphase[m] == PHASE_NEWTON)
                rbs_newton_converge(n, (int)m, a.n_part, a.max_iter, 0.01, 1, a.early_exit);
            if (n.mphase[m] != PHASE_DONE) atomicAdd(left, 1);  // integer counter
        }
        passes++;
        grid.sync();
        if (*(volatile int*)left == 0) break;  // the same value for every thread
    }
    if (gtid == 0) a.counters[3] = passes;
}
