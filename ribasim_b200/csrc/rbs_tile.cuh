// rbs_tile.cuh — the whole window of one member tile in ONE block (no grid-wide phases).
//
// The stream path runs every phase of a pass as a grid-wide kernel over the active members: each
// intermediate (levels, flows, Jacobian values, W, the Newton increment) makes a round trip
// through HBM between two launches, and the host observes a work-list count per pass.  Members
// are independent, so nothing forces the phases to be grid-wide.  Here one block owns a tile of
// TILE members (the tile of the linear solve: its LU slots live in this block's shared memory)
// and takes them through all outer steps of the window by itself: every phase is a block-stride
// loop over (entity, member-of-the-tile) items calling the same device functions as the stream
// path, phases are separated by block barriers, a block skips the halves of a pass none of its
// members is in, and it leaves when its members are done.  What a phase writes is read back by
// the same SM a few microseconds later (L1 / L2 hits instead of HBM), there is no compaction, no
// host round trip and no tail of thinly populated launches; blocks in different phases share an
// SM, so the latency-bound solve of one tile overlaps the memory phases of its neighbours.
// Per-member arithmetic, summation orders and the accept / reject logic are those of the stream
// path: results are bitwise identical.  Full-Newton mode only.
#pragma once
#include "rbs_coop.cuh"

// (entity, member) items of a phase, member fastest
#define RBS_TILE_FOR(count)                                                                    \
    for (int it_ = threadIdx.x; it_ < (int)(count) * TILE; it_ += NT)                          \
        for (int ent = it_ / TILE, mi_ = it_ % TILE, m = m_base + mi_, once_ = 1; once_; once_ = 0) \
            if (mi_ < nm)

// k_newton_update for the members of one tile: the same row partition, per-(part, row class)
// accumulation order and final summation as rbs_update_unit; `sh` holds n_part * RBS_DZ_ROWS *
// TILE partial sums
template <int TILE>
__device__ __forceinline__ void rbs_update_tile(const Net& n, const double* c, double abstol, double reltol,
                                                int rows_per_block, int n_part, int m_base, int nm,
                                                double* sh) {
    constexpr int NT = RBS_ENT_LANES * TILE / 2;
    const int items = n_part * RBS_DZ_ROWS * TILE;
    for (int it = threadIdx.x; it < items; it += NT) {
        const int mi = it % TILE, ty = (it / TILE) % RBS_DZ_ROWS, part = it / (TILE * RBS_DZ_ROWS);
        const int m = m_base + mi;
        double acc = 0.0;
        if (mi < nm && n.mphase[m] == PHASE_NEWTON) {
            const double dt = n.mh[m], inv = 1.0 / dt;
            const int r0 = part * rows_per_block, r1 = min(r0 + rows_per_block, n.n_state);
            for (int r = r0 + ty; r < r1; r += RBS_DZ_ROWS) {
                const long long o = (long long)r * n.B + m;
                double jc = 0.0;
                for (int e = n.jrow_ptr[r]; e < n.jrow_ptr[r + 1]; e++)
                    jc += n.jv[(long long)e * n.B + m] * c[(long long)n.jcol[e] * n.B + m];
                const double zo = n.z[o];
                const double b = (dt * n.du[o] - zo) * inv;
                const double dz = (jc - b) * dt;
                const double up = n.uprev[o];
                const double s = abstol + fmax(fabs(up), fabs(up + zo)) * reltol;
                const double q = dz / s;
                acc += q * q;
                n.z[o] = zo - dz;
            }
        }
        sh[it] = acc;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < n_part * TILE; it += NT) {
        const int mi = it % TILE, part = it / TILE, m = m_base + mi;
        if (mi < nm && n.mphase[m] == PHASE_NEWTON) {
            double s = 0.0;
            for (int y = 0; y < RBS_DZ_ROWS; y++) s += sh[(part * RBS_DZ_ROWS + y) * TILE + mi];
            n.norm_part[(long long)part * n.B + m] = s;
        }
    }
}

// counters[0]: number of tiles that hit the pass limit; counters[3]: most passes of any tile
template <int TILE>
__global__ void __launch_bounds__(RBS_ENT_LANES * TILE / 2, 8 / TILE) k_window_tile(Net n, EntSched es, CoopArgs a) {
    constexpr int NT = RBS_ENT_LANES * TILE / 2;
    extern __shared__ double smem[];
    __shared__ int ph_s[TILE];
    const int m_base = blockIdx.x * TILE, nm = min(TILE, n.B - m_base);
    int passes = 0;
    bool finished = false;
    for (long long pass = 0; pass < a.max_passes; pass++) {
        if (threadIdx.x < TILE) ph_s[threadIdx.x] = threadIdx.x < nm ? n.mphase[m_base + threadIdx.x] : PHASE_DONE;
        __syncthreads();
        bool any_limit = false, any_left = false;
#pragma unroll
        for (int j = 0; j < TILE; j++) {
            any_limit |= ph_s[j] == PHASE_LIMIT;
            any_left |= ph_s[j] != PHASE_DONE;
        }
        if (!any_left) { finished = true; break; }
        // ---- finalize half: members whose Newton iteration ended in the previous pass
        if (any_limit) {
            RBS_TILE_FOR(n.n_reduced)
                if (n.mphase[m] == PHASE_LIMIT) rbs_basin_item<false, 1>(n, nullptr, false, 1, ent, m);
            __syncthreads();
            RBS_TILE_FOR(n.n_state) rbs_step_limit_item(n, a.from_ts, a.alloc_total, ent, m);
            __syncthreads();
            RBS_TILE_FOR(n.n_reduced)
                if (n.mphase[m] == PHASE_LIMIT) rbs_basin_item<false, 2>(n, nullptr, false, 2, ent, m);
            __syncthreads();
            if (threadIdx.x < nm)
                rbs_step_decide_item(n, a.dt, a.h_min, a.max_halvings, a.grow_iters, a.n_steps, a.shrink_log2,
                                     m_base + threadIdx.x);
            __syncthreads();
            RBS_TILE_FOR(n.n_state + n.n_basin) rbs_step_apply_item(n, a.predictor, ent, m);
            if (n.n_dc > 0) {
                __syncthreads();  // the control state is read by the links of the Newton half
                RBS_TILE_FOR(n.n_dc) rbs_discrete_control_item(n, 0, ent, m);
            }
            __syncthreads();
            if (threadIdx.x < TILE) ph_s[threadIdx.x] = threadIdx.x < nm ? n.mphase[m_base + threadIdx.x] : PHASE_DONE;
            __syncthreads();
        }
        bool any_newton = false;
#pragma unroll
        for (int j = 0; j < TILE; j++) any_newton |= ph_s[j] == PHASE_NEWTON;
        passes++;
        if (!any_newton) continue;  // block-uniform
        // ---- Newton half: members in PHASE_NEWTON, Jacobian refreshed every iteration
        RBS_TILE_FOR(n.n_reduced)
            if (n.mphase[m] == PHASE_NEWTON) rbs_basin_item<true, 1>(n, nullptr, true, 0, ent, m);
        __syncthreads();
        RBS_TILE_FOR(n.n_conn)
            if (n.mphase[m] == PHASE_NEWTON) rbs_link_item<true>(n, ent, m, 0);
        __syncthreads();
        if (n.n_cc > 0) {
            RBS_TILE_FOR(n.n_cc)
                if (n.mphase[m] == PHASE_NEWTON) rbs_cc_item<true>(n, ent, m);
            __syncthreads();
            RBS_TILE_FOR(a.n_list_cc)
                if (n.mphase[m] == PHASE_NEWTON) rbs_link_item<true>(n, a.list_cc[ent], m, CONTROL_CONTINUOUS);
            __syncthreads();
        }
        if (n.n_pid > 0) {
            RBS_TILE_FOR(n.n_pid)
                if (n.mphase[m] == PHASE_NEWTON) rbs_pid_item<true>(n, n.ur, a.pid_fused, ent, m);
            __syncthreads();
            if (!a.pid_fused) {
                RBS_TILE_FOR(a.n_list_pid)
                    if (n.mphase[m] == PHASE_NEWTON) rbs_link_item<true>(n, a.list_pid[ent], m, CONTROL_PID);
                __syncthreads();
            }
        }
        RBS_TILE_FOR(n.n_lu + n.n_reduced)
            if (n.mphase[m] == PHASE_NEWTON) rbs_assemble_item(n, a.dsrc_ptr, a.dsrc_code, ent, m);
        __syncthreads();
        rbs_lin_tile<TILE>(n, es, smem, m_base, 0, a.c);
        __syncthreads();  // the slots double as the scratch of the update
        rbs_update_tile<TILE>(n, a.c, a.abstol, a.reltol, a.rows_per_block, a.n_part, m_base, nm, smem);
        __syncthreads();
        if (threadIdx.x < nm && n.mphase[m_base + threadIdx.x] == PHASE_NEWTON)
            rbs_newton_converge(n, m_base + threadIdx.x, a.n_part, a.max_iter, 0.01, 1, a.early_exit);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (!finished) atomicAdd(a.counters, 1);
        atomicMax(a.counters + 3, passes);
    }
}
