// rbs_fused.cuh — round-2 kernels of the step path: M ensemble members per thread.
//
// The entry-schedule solve of rbs_kernels.cuh (`rbs_ent_levels`) gives every entry of a level to
// one thread that carries two members; ~90 instructions of schedule decoding surround the three
// flops of a product, so the kernel is bound by instruction issue and by the latency of a level,
// not by bytes (profiles/r1_k_newton_linear_ent.txt: 33 % issue slots, 4 % of the HBM roofline).
// Here a thread carries M = 4 or 8 members of its entry: the schedule words are decoded once per
// M members, a product costs M/2 128-bit shared-memory loads per operand, and the M accumulation
// chains are independent (instruction-level parallelism instead of resident warps).  The slots
// of a tile are laid out S2[chunk][slot] with one 16-byte chunk = 2 members, so the 32 lanes of a
// warp, which read 32 different slots, spread over all bank groups.
// Arithmetic, summation order and team butterflies are those of `rbs_ent_levels` (and of
// symbolic.numeric_entries): results are bitwise identical to the 2-member kernel.
#pragma once
#include "rbs_kernels.cuh"

__device__ __forceinline__ void rbs_cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}

template <int M>
__device__ __forceinline__ void rbs_ent_levels_m(double2* S2, const EntSched& es, int lv0, int lv1) {
    constexpr int C = M / 2;
    constexpr int EL = RBS_ENT_LANES;  // = blockDim.x: one thread per entry lane
    const int NS = es.n_slots;
    const int lane_e = threadIdx.x;
    const int warp_lane0 = (int)(threadIdx.x & ~31u);
    const unsigned long long* __restrict__ ent = es.ent;
    const unsigned long long* __restrict__ prod = es.prod;
    for (int lv = lv0; lv < lv1; lv++) {
        const unsigned long long lw = es.lvl[lv];
        const int e0 = (int)(lw & 0xFFFFu), e1 = (int)((lw >> 16) & 0xFFFFu), tl = (int)(lw >> 32);
        if (warp_lane0 < min((e1 - e0) << tl, EL)) {
            const int ts = 1 << tl, t = lane_e & (ts - 1), stride = EL >> tl;
            for (int base = e0; base < e1; base += stride) {  // warp-uniform trip count
                const int k = base + (lane_e >> tl);
                const bool valid = k < e1;
                double acc[M];
#pragma unroll
                for (int j = 0; j < M; j++) acc[j] = 0.0;
                unsigned e = 0, q0 = 0, cf = 0;
                if (valid) {
                    const unsigned long long ew = ent[k];
                    e = (unsigned)ew & 0xFFFFu; q0 = ((unsigned)ew) >> 16; cf = (unsigned)(ew >> 32);
                    const unsigned q1 = q0 + (cf & 0xFFu);
                    for (unsigned q = q0 + t; q < q1; q += ts) {
                        const unsigned long long pw = prod[q];
                        const double2* Lp = S2 + ((unsigned)pw & 0xFFFFu);
                        const double2* Pp = S2 + (((unsigned)pw) >> 16);
                        const double2* Up = S2 + ((unsigned)(pw >> 32) & 0xFFFFu);
#pragma unroll
                        for (int c = 0; c < C; c++) {
                            const double2 l = Lp[c * NS], p = Pp[c * NS], u = Up[c * NS];
                            acc[2 * c] += (l.x * p.x) * u.x;
                            acc[2 * c + 1] += (l.y * p.y) * u.y;
                        }
                    }
                }
                if (tl) {  // uniform; teams are aligned groups of lanes
#pragma unroll
                    for (int j = 0; j < M; j++) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
                    if (tl > 1) {
#pragma unroll
                        for (int j = 0; j < M; j++) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
                        if (tl > 2) {
#pragma unroll
                            for (int j = 0; j < M; j++) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
                        }
                    }
                }
                if (valid && t == 0) {
                    // the reciprocal pivot: p of the first product, or q0 itself without products
                    const unsigned r = (cf & RBS_ENT_SCALE) ? ((cf & 0xFFu) ? (((unsigned)prod[q0]) >> 16) : q0) : 0u;
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        double2* T = S2 + c * NS + e;
                        double2 v = *T;
                        if (cf & RBS_ENT_FRESH) { v.x = 0.0; v.y = 0.0; }
                        if (cf & RBS_ENT_SCALE) {
                            const double2 rv = S2[c * NS + r];
                            v.x *= rv.x; v.y *= rv.y;
                        }
                        v.x -= acc[2 * c]; v.y -= acc[2 * c + 1];
                        if (cf & RBS_ENT_RCP) { v.x = 1.0 / v.x; v.y = 1.0 / v.y; }
                        *T = v;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// Members of the tile [m0, m0 + M) of the launch domain that are in the Newton phase (-1 = idle),
// and per 16-byte chunk whether its two members are adjacent and aligned in memory.
template <int M>
__device__ __forceinline__ void rbs_tile_members(const Net& n, int m0, int* mem_s, int* pair_s) {
    if (threadIdx.x < M) {
        const int j = threadIdx.x;
        int mm = m0 + j < n.NA ? RBS_MEMBER(n, m0 + j) : -1;
        if (mm >= 0 && n.mphase[mm] != PHASE_NEWTON) mm = -1;
        mem_s[j] = mm;
    }
    __syncthreads();
    if (threadIdx.x < M / 2) {
        const int a = mem_s[2 * threadIdx.x], b = mem_s[2 * threadIdx.x + 1];
        pair_s[threadIdx.x] = a >= 0 && b == a + 1 && ((a | n.B) & 1) == 0;
    }
    __syncthreads();
}

// Newton linear solve of one M-member tile per block on the compact entry schedule (full-Newton
// path: every call refactors, the factors are not kept).  Input: the assembled slot values in
// `lu` and the reduced right-hand side in `br` (k_assemble_linear); output: x = W^-1 b_r.
template <int M>
__global__ void __launch_bounds__(RBS_ENT_LANES) k_newton_linear_m(Net n, EntSched es, double* __restrict__ x) {
    RBS_PDL();
    constexpr int C = M / 2, NT = RBS_ENT_LANES;
    extern __shared__ double smem[];
    double2* S2 = (double2*)smem;
    __shared__ int mem_s[M], pair_s[M / 2];
    const int NS = es.n_slots;
    const int m0 = blockIdx.x * M;
    if (m0 >= n.NA) return;
    rbs_tile_members<M>(n, m0, mem_s, pair_s);
    if (es.smem_sched) {
        unsigned long long* sw = (unsigned long long*)(S2 + (size_t)NS * C);
        for (int w = threadIdx.x; w < es.n_words; w += NT) sw[w] = es.lvl[w];
        es.ent = sw + (es.ent - es.lvl);
        es.prod = sw + (es.prod - es.lvl);
        es.lvl = sw;
    }
    const long long B = n.B;
    const uint16_t* __restrict__ perm = es.perm;
    {
        const int items = (es.n_load + n.n_reduced) * C;
        for (int it = threadIdx.x; it < items; it += NT) {
            const int c = it % C, k = it / C;
            const double* src;
            int slot;
            if (k < es.n_load) {
                const unsigned w = es.load[k];
                src = n.lu + (long long)(w & 0xFFFFu) * B;
                slot = (int)(w >> 16);
            } else {
                const int i = k - es.n_load;
                src = n.br + (long long)(int)perm[i] * B;
                slot = es.y0 + i;
            }
            double2* dst = S2 + c * NS + slot;
            const int ma = mem_s[2 * c], mb = mem_s[2 * c + 1];
            if (pair_s[c]) rbs_cp_async16(dst, src + ma);
            else {
                if (ma >= 0) rbs_cp_async8(&dst->x, src + ma);
                if (mb >= 0) rbs_cp_async8(&dst->y, src + mb);
            }
        }
    }
    rbs_cp_async_wait_all();
    __syncthreads();
    rbs_ent_levels_m<M>(S2, es, 0, es.n_elim + es.n_bwd);
    {
        const int items = n.n_reduced * C;
        for (int it = threadIdx.x; it < items; it += NT) {
            const int c = it % C, i = it / C;
            const double2 v = S2[c * NS + es.y0 + i];
            double* dst = x + (long long)(int)perm[i] * B;
            const int ma = mem_s[2 * c], mb = mem_s[2 * c + 1];
            if (pair_s[c]) *(double2*)(dst + ma) = v;
            else {
                if (ma >= 0) dst[ma] = v.x;
                if (mb >= 0) dst[mb] = v.y;
            }
        }
    }
}

// ---- end of a pass: convergence test + the whole finalize half in ONE kernel ---------------------
// The stream path of round 1 ended a Newton iteration with the convergence test inside the
// single-block k_compact and ran the finalize half (limit_flow! on the proposed state, the
// negative-storage check, accept / reject, the exact-forcing bookkeeping, DiscreteControl) as five
// to six grid-wide launches on a compacted list of the members whose iteration had ended.  Those
// are a quarter of the members, scattered, so every launch touched most sectors of the state
// arrays for a quarter of useful lanes, and each phase boundary was a launch.
// Here a thread-block CLUSTER owns a tile of 32 consecutive members (lane = member: every access
// is one coalesced 256-byte row), its CL blocks split the entities, and the phases are separated
// by cluster barriers (barrier.cluster, release / acquire) instead of launches; what a phase wrote
// is re-read from L2 by the same SMs a microsecond later.  Clusters without a member whose
// iteration ended leave after the convergence test.  The per-member arithmetic is that of the
// round-1 kernels (same device functions): results are bitwise identical.
#include <cooperative_groups.h>

struct FinArgs {
    const int* from_ts;
    const double* alloc_total;
    double dt, h_min, kappa;
    int max_halvings, grow_iters, n_steps, shrink_log2, predictor;
    int n_part, max_iter, full_newton, early_exit;
    int m_lo, m_hi;
};

#define RBS_FIN_THREADS 512

template <int CL>
__device__ __forceinline__ void rbs_fin_sync() {
    if (CL > 1) cooperative_groups::this_cluster().sync();
    else __syncthreads();
}

template <int CL>
__global__ void __launch_bounds__(RBS_FIN_THREADS) k_finalize(Net n, FinArgs a) {
    RBS_PDL();
    const unsigned FULL = 0xffffffffu;
    const int rank = CL > 1 ? (int)cooperative_groups::this_cluster().block_rank() : 0;
    const int tile = blockIdx.x / CL;
    const int lane = threadIdx.x & 31, W = RBS_FIN_THREADS >> 5;
    const int gw = rank * W + (threadIdx.x >> 5), GW = CL * W;
    const int m = a.m_lo + tile * 32 + lane;
    const bool in = m < a.m_hi;
    // every member in the Newton phase iterated in this pass: nlsolve!'s convergence test
    if (gw == 0 && in && n.mphase[m] == PHASE_NEWTON)
        rbs_newton_converge(n, m, a.n_part, a.max_iter, a.kappa, a.full_newton, a.early_exit);
    rbs_fin_sync<CL>();
    const bool lim = in && n.mphase[m] == PHASE_LIMIT;
    if (__ballot_sync(FULL, lim) == 0u) return;  // the same 32 members in every warp of the cluster
    // storage / level at the proposed state uprev + z, negative-storage flag
    for (int ent = gw; ent < n.n_reduced; ent += GW)
        if (lim) rbs_basin_item<false, 1>(n, nullptr, false, 1, ent, m);
    rbs_fin_sync<CL>();
    // limit_flow!: u = clamp(uprev + z)
    for (int s = gw; s < n.n_state; s += GW)
        if (lim) rbs_step_limit_item(n, a.from_ts, a.alloc_total, s, m);
    rbs_fin_sync<CL>();
    // members the limiter changed: storage of the clamped state decides about negativity
    const bool cl = lim && n.mclamp[m] != 0;
    if (__ballot_sync(FULL, cl) != 0u) {
        for (int ent = gw; ent < n.n_reduced; ent += GW)
            if (cl) rbs_basin_item<false, 2>(n, nullptr, false, 2, ent, m);
        rbs_fin_sync<CL>();
    }
    if (gw == 0 && lim)
        rbs_step_decide_item(n, a.dt, a.h_min, a.max_halvings, a.grow_iters, a.n_steps, a.shrink_log2, m);
    rbs_fin_sync<CL>();
    for (int ent = gw; ent < n.n_state + n.n_basin; ent += GW)
        if (lim) rbs_step_apply_item(n, a.predictor, ent, m);
    if (n.n_dc > 0) {
        rbs_fin_sync<CL>();  // the control state is read by the links of the next Newton iteration
        for (int i = gw; i < n.n_dc; i += GW)
            if (lim) rbs_discrete_control_item(n, 0, i, m);
    }
}

// Work list of a member group for the next pass: the members of every aligned tile of 8 that
// still has a member to advance (tiles stay whole so that the M-member solve finds aligned,
// adjacent members; finished members of a listed tile are skipped by their phase).  One block;
// counts[0] = list length, counts[2] = outer steps completed by every member of the group.
__global__ void k_compact_tiles(Net n, int m_lo, int m_hi, int* __restrict__ alist_out,
                                int* __restrict__ counts) {
    RBS_PDL();
    __shared__ int sa[1024];
    const int t = threadIdx.x, T = blockDim.x;
    const int nm = m_hi - m_lo;
    const int per = ((nm + T - 1) / T + 7) & ~7;
    const int lo = m_lo + min(t * per, nm), hi = min(lo + per, m_hi);
    int ca = 0, mn = 0x7fffffff;
    for (int m8 = lo; m8 < hi; m8 += 8) {
        const int e8 = min(m8 + 8, hi);
        bool any = false;
        for (int m = m8; m < e8; m++) {
            any |= n.mphase[m] != PHASE_DONE;
            mn = min(mn, n.mstep[m]);
        }
        if (any) ca += e8 - m8;
    }
    sa[t] = ca;
    __syncthreads();
    for (int off = 1; off < T; off <<= 1) {  // inclusive Hillis-Steele scan
        const int va = t >= off ? sa[t - off] : 0;
        __syncthreads();
        sa[t] += va;
        __syncthreads();
    }
    int pa = sa[t] - ca;
    const int total = sa[T - 1];
    __syncthreads();
    for (int m8 = lo; m8 < hi; m8 += 8) {
        const int e8 = min(m8 + 8, hi);
        bool any = false;
        for (int m = m8; m < e8; m++) any |= n.mphase[m] != PHASE_DONE;
        if (any) for (int m = m8; m < e8; m++) alist_out[pa++] = m;
    }
    sa[t] = mn;
    __syncthreads();
    for (int off = T >> 1; off > 0; off >>= 1) {
        if (t < off) sa[t] = min(sa[t], sa[t + off]);
        __syncthreads();
    }
    if (t == 0) { counts[0] = total; counts[1] = 0; counts[2] = sa[0]; }
}

// ---- work lists of the next pass, device-resident counts ----------------------------------------
// Ordered compaction of the members [m_lo, m_hi) (integer work only; the convergence test has
// moved into k_newton_update): `alist_out` = members not DONE, `llist_out` = members in
// PHASE_LIMIT, counts[0..1] = their lengths, counts[2] = outer steps completed by every member of
// the group, counts[3] = passes so far.  Warp-shuffle scans: ~10 us against ~36 us of k_compact.
__device__ __forceinline__ int rbs_block_exscan(int v, int* warp_tot, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
        int t = lane < nw ? warp_tot[lane] : 0;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, t, off);
            if (lane >= off) t += y;
        }
        warp_tot[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    total = warp_tot[nw - 1];
    const int base = w ? warp_tot[w - 1] : 0;
    __syncthreads();
    return base + x - v;
}

__global__ void __launch_bounds__(1024) k_compact2(Net n, int m_lo, int m_hi, int* __restrict__ alist_out,
                                                   int* __restrict__ llist_out, int* __restrict__ counts) {
    RBS_PDL();
    __shared__ int wt[32];
    const int t = threadIdx.x, T = blockDim.x;
    const int nm = m_hi - m_lo;
    const int per = (nm + T - 1) / T;
    const int lo = m_lo + min(t * per, nm), hi = min(lo + per, m_hi);
    int ca = 0, cl = 0, mn = 0x7fffffff;
    for (int m = lo; m < hi; m++) {
        const int ph = n.mphase[m];
        ca += ph != PHASE_DONE;
        cl += ph == PHASE_LIMIT;
        mn = min(mn, n.mstep[m]);
    }
    int ta, tl;
    int pa = rbs_block_exscan(ca, wt, ta);
    int pl = rbs_block_exscan(cl, wt, tl);
    for (int m = lo; m < hi; m++) {
        const int ph = n.mphase[m];
        if (ph != PHASE_DONE) alist_out[pa++] = m;
        if (ph == PHASE_LIMIT) llist_out[pl++] = m;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    if ((t & 31) == 0) wt[t >> 5] = mn;
    __syncthreads();
    if (t == 0) {
        for (int w = 1; w < (T + 31) / 32; w++) mn = min(mn, wt[w]);
        const int passes = counts[3] + 1;
        counts[0] = ta; counts[1] = tl; counts[2] = mn; counts[3] = passes;
    }
}


// ---- finalize half: storage of the clamped state + accept / reject in one launch -----------------
// k_basin<false, 2> (members the limiter changed) followed by k_step_decide: the decision needs the
// negative-storage flags of all basins, so the last block to finish (a counter) takes it for the
// members of the list.  Same device functions, same results; one launch less per pass.
__global__ void __launch_bounds__(256) k_limit_post_decide(Net n, Pred pred, int* __restrict__ counter, double dt,
                                                           double h_min, int max_halvings, int grow_iters,
                                                           int n_steps, int shrink_log2) {
    RBS_PDL();
    __shared__ int last;
    const long long tid = RBS_TID;
    const int b = (int)(tid / n.NA);
    if (b < n.n_reduced) {
        const int m = RBS_MEMBER(n, (int)(tid - (long long)b * n.NA));
        if (!rbs_skip<false>(pred, m)) rbs_basin_item<false, 2>(n, nullptr, false, 2, b, m);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(counter, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *counter = 0;
    __threadfence();
    for (int j = threadIdx.x; j < n.NA; j += blockDim.x)
        rbs_step_decide_item(n, dt, h_min, max_halvings, grow_iters, n_steps, shrink_log2, RBS_MEMBER(n, j));
}
