// Newton linear solve, one ensemble member per lane.
//
// Same entry schedule as `rbs_ent_levels` (rbs_kernels.cuh; ribasim_b200/symbolic.py build_entries,
// plain slot layout) and the same arithmetic in the same order, so the results are bitwise those of
// the shared-memory kernel — but the mapping is turned around: a block owns 32 members of the launch
// domain, lane = member, and a warp owns one entry of a level at a time.  Every index of the
// schedule is warp-uniform (no per-lane decoding, no shuffles, no idle lanes in the narrow levels at
// the top of the elimination tree), every operand access is one coalesced 256-byte request.
//
// Storage.  The factor slots live in global memory (`lu`, [slot][member]; the rows behind n_lu hold
// the right-hand-side column): 23 KB per member that stay in L2 between the levels.  The reciprocal
// pivots — the middle operand of every product `(l * p) * u` and the scale of every backward entry
// — live in shared memory ([pivot][lane], 130 KB for the 500-basin network), which takes a third
// of the operand loads off L2.
//
// Schedule.  `rbs_create` unrolls the entry schedule into one linear tape of 16-byte records per
// warp (level header, entry header, products with ready-made element offsets `slot * B`): a warp
// streams through its tape with 128-bit loads (lines prefetched into L1 ahead of the reader),
// nothing is searched or decoded at run time.  A warp's entries of a level are sorted into runs of
// one kind and product count, so a run's loop body is straight-line code (~11 instructions per
// product, ~35 per two-product entry).  Measured: a block is bound by the operand
// round trips to L2 and the level barrier, not by instruction issue.
//
// Teams.  A team level of the entry schedule (2^tl lanes share an entry's product list: lane t
// sums the products q0 + t, q0 + t + 2^tl, ...; the partial sums are combined by a shuffle
// butterfly) is reproduced either inside a warp — 2^tl accumulators per lane, combined in the
// butterfly's order (levels with enough entries to keep every warp busy) — or, in the narrow
// levels at the top of the elimination tree, by a team of 2^tl WARPS that exchange their partial
// sums through shared memory, so that a long product list is not one warp's chain of dependent
// memory round trips.
//
// Replaces: the KLU refactor + solve of core/src/differentiation.jl:321-345 (same role as
// k_newton_linear_ent).
#pragma once

struct LaneTape {
    const uint4* rec;        // records of all warps
    const int* start;        // first record of warp w
    const unsigned* piv_off; // element offset (slot * B) of pivot i in `lu`
    int n_lv, n_piv;         // levels, pivots
    const int* iperm;        // reduced row -> position in the elimination order
};
// record layouts
//   level header : x = runs of this warp (team level: 1 if the warp has a unit), y = 1 for a team
//                  level, z = tl, w = team member t of this warp's unit
//   run header   : x = kind (RBS_LANE_K_*), y = products per entry (0..4) or RBS_LANE_DYN (every
//                  entry header carries its own count), z = entries
//   entry header : x = element offset of the target slot, y = byte offset of its pivot row in shared
//                  memory (RCP: the pivot this entry produces, SCALE: the pivot it is scaled by),
//                  z = product count | flags, w = element offset in x (BWD)
//   product      : x = element offset of l, y = byte offset of the pivot row, z = element offset of u
// A warp's entries of a level are independent, so the host sorts them into runs of one kind and —
// in levels without teams — one product count: the loop body of a run is straight-line code (no
// flag tests, no count dispatch, constant tape stride).
#define RBS_LANE_K_RCP 0u   // pivot: v = 1 / (old - acc) -> shared memory (+ `lu` for members that refactor)
#define RBS_LANE_K_FAC 1u   // factor slot: v = old - acc, members that refactor
#define RBS_LANE_K_YF 2u    // right-hand-side slot in the elimination (fused forward substitution)
#define RBS_LANE_K_YB 3u    // backward substitution: v = old * pivot - acc, final: also stored to x
#define RBS_LANE_DYN 0xFFu

// address of element `off` of a member column: one IMAD.WIDE (written as PTX so that the compiler
// does not re-derive the column base from its parts for every operand)
__device__ __forceinline__ double* rbs_lane_at(const double* Lm, unsigned off) {
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(a) : "r"(off), "l"((unsigned long long)Lm));
    double* p = (double*)a;
    __builtin_assume(__isGlobal(p));
    return p;
}

// the tape is read once per block, front to back: its lines are pulled into L1 ahead of the reader
__device__ __forceinline__ void rbs_lane_prefetch(const uint4* p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// N products at tp: loads in flight together, then a[(BASE + j) % TS] += (l * p) * u in list order
template <int N, int BASE, int TS>
__device__ __forceinline__ void rbs_lane_chunk(const uint4* __restrict__ tp, double (&a)[TS],
                                               const double* __restrict__ Lm, const char* piv) {
    uint4 r[N];
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = __ldg(tp + j);
    double l[N], p[N], u[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        l[j] = *rbs_lane_at(Lm, r[j].x);
        u[j] = *rbs_lane_at(Lm, r[j].z);
        p[j] = *(const double*)(piv + r[j].y);
    }
#pragma unroll
    for (int j = 0; j < N; j++) a[(BASE + j) % TS] = a[(BASE + j) % TS] + (l[j] * p[j]) * u[j];
}
template <int BASE, int TS>
__device__ __forceinline__ void rbs_lane_chunk_n(int n_c, const uint4* __restrict__ tp, double (&a)[TS],
                                                 const double* __restrict__ Lm, const char* piv) {
    if (n_c >= 4) rbs_lane_chunk<4, BASE, TS>(tp, a, Lm, piv);
    else if (n_c == 3) rbs_lane_chunk<3, BASE, TS>(tp, a, Lm, piv);
    else if (n_c == 2) rbs_lane_chunk<2, BASE, TS>(tp, a, Lm, piv);
    else rbs_lane_chunk<1, BASE, TS>(tp, a, Lm, piv);
}
// the cnt products at tp with TS interleaved partial sums, combined in the butterfly's order
template <int TS>
__device__ __forceinline__ double rbs_lane_products(const uint4* __restrict__ tp, int cnt,
                                                    const double* __restrict__ Lm, const char* piv) {
    double a[TS];
#pragma unroll
    for (int j = 0; j < TS; j++) a[j] = 0.0;
    if (TS == 8) {
        for (int c = 0; c < cnt; c += 8) {
            rbs_lane_chunk_n<0, TS>(cnt - c, tp + c, a, Lm, piv);
            if (c + 4 < cnt) rbs_lane_chunk_n<4, TS>(cnt - c - 4, tp + c + 4, a, Lm, piv);
        }
    } else {
        for (int c = 0; c < cnt; c += 4) rbs_lane_chunk_n<0, TS>(cnt - c, tp + c, a, Lm, piv);
    }
#pragma unroll
    for (int s = 1; s < TS; s <<= 1)
#pragma unroll
        for (int j = 0; j < TS; j += 2 * s) a[j] += a[j + s];
    return a[0];
}

template <int V> struct RbsInt { static constexpr int value = V; };

// per-thread context of the lane kernel
struct LaneCtx {
    double* Lm;        // this lane's member column of `lu`
    double* xm;        // ... of the solution vector
    char* piv;         // ... of the pivots in shared memory
    bool live, fac;    // lane has a member in the Newton phase / that member refactors
};

template <int KIND>
__device__ __forceinline__ void rbs_lane_finish(const LaneCtx& c, const uint4 eh, double o, double acc) {
    if (KIND == (int)RBS_LANE_K_RCP) {
        const double v = 1.0 / (o - acc);
        if (c.fac || !c.live) *(double*)(c.piv + eh.y) = v;
        if (c.fac) *rbs_lane_at(c.Lm, eh.x) = v;
    } else if (KIND == (int)RBS_LANE_K_FAC) {
        if (c.fac) *rbs_lane_at(c.Lm, eh.x) = o - acc;
    } else if (KIND == (int)RBS_LANE_K_YF) {
        if (c.live) *rbs_lane_at(c.Lm, eh.x) = o - acc;
    } else {
        const double v = o * *(const double*)(c.piv + eh.y) - acc;
        if (c.live) {
            *rbs_lane_at(c.Lm, eh.x) = v;
            *rbs_lane_at(c.xm, eh.w) = v;
        }
    }
}

// two entries of CNT products each at once: the loads of both are in flight together (an entry is a
// chain of dependent round trips — header, operands, reciprocal — and a warp has nothing else to
// issue meanwhile)
template <int KIND, int CNT>
__device__ __forceinline__ void rbs_lane_pair(const LaneCtx& c, const uint4* __restrict__ tp) {
    constexpr int NP = CNT > 0 ? CNT : 1;
    const uint4 e0 = __ldg(tp), e1 = __ldg(tp + 1 + CNT);
    uint4 r0[NP], r1[NP];
#pragma unroll
    for (int j = 0; j < CNT; j++) { r0[j] = __ldg(tp + 1 + j); r1[j] = __ldg(tp + 2 + CNT + j); }
    rbs_lane_prefetch(tp + 16);
    const double o0 = *rbs_lane_at(c.Lm, e0.x), o1 = *rbs_lane_at(c.Lm, e1.x);
    double l0[NP], u0[NP], p0[NP], l1[NP], u1[NP], p1[NP];
#pragma unroll
    for (int j = 0; j < CNT; j++) {
        l0[j] = *rbs_lane_at(c.Lm, r0[j].x); u0[j] = *rbs_lane_at(c.Lm, r0[j].z);
        l1[j] = *rbs_lane_at(c.Lm, r1[j].x); u1[j] = *rbs_lane_at(c.Lm, r1[j].z);
        p0[j] = *(const double*)(c.piv + r0[j].y); p1[j] = *(const double*)(c.piv + r1[j].y);
    }
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int j = 0; j < CNT; j++) { a0 = a0 + (l0[j] * p0[j]) * u0[j]; a1 = a1 + (l1[j] * p1[j]) * u1[j]; }
    rbs_lane_finish<KIND>(c, e0, o0, a0);
    rbs_lane_finish<KIND>(c, e1, o1, a1);
}

// one run of n entries of one kind; CNT >= 0: every entry has CNT products (levels without teams)
template <int KIND, int CNT, int TS>
__device__ __forceinline__ const uint4* rbs_lane_run(const LaneCtx& c, const uint4* __restrict__ tp, int n) {
    if (CNT >= 0 && CNT <= 2) {
#pragma unroll 1
        for (; n >= 2; n -= 2) {
            rbs_lane_pair<KIND, (CNT >= 0 && CNT <= 2 ? CNT : 0)>(c, tp);
            tp += 2 * (1 + CNT);
        }
    }
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const uint4 eh = __ldg(tp);
        rbs_lane_prefetch(tp + 16);
        const double o = *rbs_lane_at(c.Lm, eh.x);
        double acc = 0.0;
        if (CNT >= 0) {
            if (CNT > 0) {
                double a[1] = {0.0};
                rbs_lane_chunk<(CNT > 0 ? CNT : 1), 0, 1>(tp + 1, a, c.Lm, c.piv);
                acc = a[0];
            }
            tp += 1 + CNT;
        } else {
            const int cnt = (int)(eh.z & 0xFFu);
            acc = rbs_lane_products<TS>(tp + 1, cnt, c.Lm, c.piv);
            tp += 1 + cnt;
        }
        rbs_lane_finish<KIND>(c, eh, o, acc);
    }
    return tp;
}
template <int KIND, int TS>
__device__ __forceinline__ const uint4* rbs_lane_run_kind(const LaneCtx& c, const uint4* __restrict__ tp,
                                                          unsigned cnt_code, int n) {
    if (TS == 1) {
        switch (cnt_code) {
        case 0: return rbs_lane_run<KIND, 0, 1>(c, tp, n);
        case 1: return rbs_lane_run<KIND, 1, 1>(c, tp, n);
        case 2: return rbs_lane_run<KIND, 2, 1>(c, tp, n);
        case 3: return rbs_lane_run<KIND, 3, 1>(c, tp, n);
        case 4: return rbs_lane_run<KIND, 4, 1>(c, tp, n);
        default: break;
        }
    }
    return rbs_lane_run<KIND, -1, TS>(c, tp, n);
}
// the runs of this warp in a level without warp teams
template <int TS>
__device__ __forceinline__ const uint4* rbs_lane_level(const LaneCtx& c, const uint4* __restrict__ tp, int n_runs) {
    for (int r = 0; r < n_runs; r++) {
        const uint4 rh = __ldg(tp);
        tp++;
        if (rh.x == RBS_LANE_K_RCP) tp = rbs_lane_run_kind<(int)RBS_LANE_K_RCP, TS>(c, tp, rh.y, (int)rh.z);
        else if (rh.x == RBS_LANE_K_FAC) tp = rbs_lane_run_kind<(int)RBS_LANE_K_FAC, TS>(c, tp, rh.y, (int)rh.z);
        else if (rh.x == RBS_LANE_K_YF) tp = rbs_lane_run_kind<(int)RBS_LANE_K_YF, TS>(c, tp, rh.y, (int)rh.z);
        else tp = rbs_lane_run_kind<(int)RBS_LANE_K_YB, TS>(c, tp, rh.y, (int)rh.z);
    }
    return tp;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32)
k_newton_linear_lane(Net n, LaneTape tape, double* __restrict__ x) {
    RBS_PDL();
    extern __shared__ double lane_smem[];
    double* part_s = lane_smem;                            // [NW][32] team partial sums
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    int m = j < n.NA ? RBS_MEMBER(n, j) : -1;
    if (m >= 0 && n.mphase[m] != PHASE_NEWTON) m = -1;
    LaneCtx c;
    c.live = m >= 0;
    c.fac = c.live && n.mjac[m] != 0;
    // lanes without a member compute on the columns of the block's first member and store nothing
    const int ma = c.live ? m : RBS_MEMBER(n, blockIdx.x * 32);
    c.Lm = n.lu + ma;
    c.xm = x + ma;
    c.piv = (char*)(lane_smem + NW * 32) + lane * 8;       // [n_piv][32] reciprocal pivots
    LIN_CLK(0);
    // the right-hand-side column is already behind the factor slots (k_assemble_linear, Net::lane_iperm);
    // members that keep their factors need their stored pivots
    if (__syncthreads_or(c.live && !c.fac))
        for (int i = w; i < tape.n_piv; i += NW) *(double*)(c.piv + i * 256) = *rbs_lane_at(c.Lm, tape.piv_off[i]);
    __syncthreads();
    LIN_CLK(2);
    const uint4* __restrict__ tp = tape.rec + tape.start[w];
    rbs_lane_prefetch(tp + 8);
    for (int lv = 0; lv < tape.n_lv; lv++) {
        const uint4 lh = __ldg(tp);
        tp++;
        if (lh.y) {
            // narrow level: one unit (entry, team member) per warp, partial sums through shared memory
            uint4 eh = make_uint4(0, 0, 0, 0);
            double o = 0.0;
            if (lh.x) {
                eh = __ldg(tp);
                rbs_lane_prefetch(tp + 16);
                const int cnt = (int)(eh.z & 0xFFu);
                if (lh.w == 0) o = *rbs_lane_at(c.Lm, eh.x);
                part_s[w * 32 + lane] = rbs_lane_products<1>(tp + 1, cnt, c.Lm, c.piv);
                tp += 1 + cnt;
            }
            __syncthreads();
            if (lh.x && lh.w == 0) {
                const int tl = (int)lh.z;
                double a[8];
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = i < (1 << tl) ? part_s[(w + i) * 32 + lane] : 0.0;
                // butterfly of the shuffle version: lane t adds lane t ^ 1, then t ^ 2, t ^ 4
                a[0] += a[1]; a[2] += a[3]; a[4] += a[5]; a[6] += a[7];
                if (tl >= 2) { a[0] += a[2]; a[4] += a[6]; }
                if (tl >= 3) { a[0] += a[4]; }
                const unsigned kind = eh.z >> 16;
                if (kind == RBS_LANE_K_RCP) rbs_lane_finish<(int)RBS_LANE_K_RCP>(c, eh, o, a[0]);
                else if (kind == RBS_LANE_K_FAC) rbs_lane_finish<(int)RBS_LANE_K_FAC>(c, eh, o, a[0]);
                else if (kind == RBS_LANE_K_YF) rbs_lane_finish<(int)RBS_LANE_K_YF>(c, eh, o, a[0]);
                else rbs_lane_finish<(int)RBS_LANE_K_YB>(c, eh, o, a[0]);
            }
        } else {
            const int tl = (int)lh.z;
            if (tl == 0) tp = rbs_lane_level<1>(c, tp, (int)lh.x);
            else if (tl == 1) tp = rbs_lane_level<2>(c, tp, (int)lh.x);
            else if (tl == 2) tp = rbs_lane_level<4>(c, tp, (int)lh.x);
            else tp = rbs_lane_level<8>(c, tp, (int)lh.x);
        }
        __syncthreads();
#ifdef RBS_LIN_PROFILE
        if (blockIdx.x == 0 && threadIdx.x == 0 && lv < 64) g_lin_lvl[lv] = clock64();
#endif
    }
    LIN_CLK(7);
}
