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
// streams through its tape with 128-bit loads, nothing is searched or decoded at run time.  The
// kernel is bound by instruction issue (one block per SM, 32 members), so the tape exists to keep
// the instruction count per product at ~11.
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

#define RBS_LANE_RCP 0x100u     // as RBS_ENT_RCP / RBS_ENT_SCALE
#define RBS_LANE_SCALE 0x200u
#define RBS_LANE_Y 0x800u       // target is a right-hand-side slot (every member updates it)
#define RBS_LANE_BWD 0x1000u    // backward-substitution entry: the value is final, also stored to x

struct LaneTape {
    const uint4* rec;        // records of all warps
    const int* start;        // first record of warp w
    const unsigned* piv_off; // element offset (slot * B) of pivot i in `lu`
    int n_lv, n_piv;         // levels, pivots
    const int* iperm;        // reduced row -> position in the elimination order
};
// record layouts
//   level header : x = entries of this warp (team level: 1 if the warp has a unit), y = 1 for a team
//                  level, z = tl, w = team member t of this warp's unit
//   entry header : x = element offset of the target slot, y = byte offset of its pivot row in shared
//                  memory (RCP: the pivot this entry produces, SCALE: the pivot it is scaled by),
//                  z = product count of this record run | flags, w = element offset in x (BWD)
//   product      : x = element offset of l, y = byte offset of the pivot row, z = element offset of u

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

template <int NW>
__global__ void __launch_bounds__(NW * 32)
k_newton_linear_lane(Net n, LaneTape tape, double* __restrict__ x) {
    RBS_PDL();
    extern __shared__ double lane_smem[];
    double* part_s = lane_smem;                            // [NW][32] team partial sums
    char* piv = (char*)(lane_smem + NW * 32);              // [n_piv][32] reciprocal pivots
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    int m = j < n.NA ? RBS_MEMBER(n, j) : -1;
    if (m >= 0 && n.mphase[m] != PHASE_NEWTON) m = -1;
    const bool live = m >= 0;
    const bool fac = live && n.mjac[m] != 0;
    // lanes without a member compute on the columns of the block's first member and store nothing
    const int ma = live ? m : RBS_MEMBER(n, blockIdx.x * 32);
    double* Lm = n.lu + ma;
    double* xm = x + ma;
    piv += lane * 8;
    LIN_CLK(0);
    // the right-hand-side column is already behind the factor slots (k_assemble_linear, Net::lane_iperm);
    // members that keep their factors need their stored pivots
    if (__syncthreads_or(live && !fac))
        for (int i = w; i < tape.n_piv; i += NW) *(double*)(piv + i * 256) = *rbs_lane_at(Lm, tape.piv_off[i]);
    __syncthreads();
    LIN_CLK(2);
    const uint4* __restrict__ tp = tape.rec + tape.start[w];
    uint4 cur = __ldg(tp);
    rbs_lane_prefetch(tp + 8);
    rbs_lane_prefetch(tp + 16);
    auto finish = [&](const uint4 eh, double o, double acc) {
        double v = o;
        if (eh.z & RBS_LANE_SCALE) v *= *(const double*)(piv + eh.y);
        v -= acc;
        if (eh.z & RBS_LANE_Y) {
            if (live) {
                *rbs_lane_at(Lm, eh.x) = v;
                if (eh.z & RBS_LANE_BWD) *rbs_lane_at(xm, eh.w) = v;
            }
        } else {
            if (eh.z & RBS_LANE_RCP) {
                v = 1.0 / v;
                if (fac || !live) *(double*)(piv + eh.y) = v;
            }
            if (fac) *rbs_lane_at(Lm, eh.x) = v;
        }
    };
    for (int lv = 0; lv < tape.n_lv; lv++) {
        const uint4 lh = cur;   // level header; `cur` is always the record at tp, loaded one step ahead
        tp++;
        cur = __ldg(tp);
        if (lh.y) {
            // narrow level: one unit (entry, team member) per warp, partial sums through shared memory
            uint4 eh = make_uint4(0, 0, 0, 0);
            double o = 0.0;
            if (lh.x) {
                eh = cur;
                const int cnt = (int)(eh.z & 0xFFu);
                const uint4* tn = tp + 1 + cnt;
                cur = __ldg(tn);
                rbs_lane_prefetch(tn + 16);
                if (lh.w == 0) o = *rbs_lane_at(Lm, eh.x);
                part_s[w * 32 + lane] = rbs_lane_products<1>(tp + 1, cnt, Lm, piv);
                tp = tn;
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
                finish(eh, o, a[0]);
            }
        } else {
            // the entries of this warp, their whole product lists inside the warp (the switch on the
            // level's team size stays outside the entry loop)
            auto entries = [&](auto ts_tag) {
                constexpr int TS = decltype(ts_tag)::value;
                for (int i = 0; i < (int)lh.x; i++) {
                    const uint4 eh = cur;
                    const int cnt = (int)(eh.z & 0xFFu);
                    const uint4* tn = tp + 1 + cnt;
                    cur = __ldg(tn);   // next header: its round trip overlaps this entry's operands
                    rbs_lane_prefetch(tn + 16);
                    const double o = *rbs_lane_at(Lm, eh.x);
                    const double acc = rbs_lane_products<TS>(tp + 1, cnt, Lm, piv);
                    tp = tn;
                    finish(eh, o, acc);
                }
            };
            const int tl = (int)lh.z;
            if (tl == 0) entries(RbsInt<1>{});
            else if (tl == 1) entries(RbsInt<2>{});
            else if (tl == 2) entries(RbsInt<4>{});
            else entries(RbsInt<8>{});
        }
        __syncthreads();
#ifdef RBS_LIN_PROFILE
        if (blockIdx.x == 0 && threadIdx.x == 0 && lv < 64) g_lin_lvl[lv] = clock64();
#endif
    }
    LIN_CLK(7);
}
