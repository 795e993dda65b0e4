"""Host symbolic factorisation for the batched sparse LU of `W_inner`.

The reference factorises `W_inner` with KLU (PureKLU 1.4.1 via LinearSolve,
core/src/config.jl:420-422, core/src/differentiation.jl:278-359) once per Jacobian refresh.
Here the pattern is fixed for all ensemble members, so everything that depends only on the
pattern is computed once on the host:

* a fill-reducing ordering by *rake-and-compress rounds*: every round eliminates an
  independent set of low-degree vertices (lowest degree first).  Leaves are
  raked, chain vertices are compressed, so trees and chains finish in O(log n) rounds with
  O(n) fill (a tridiagonal chain becomes cyclic reduction);
* the static fill pattern of L+U on the symmetrised pattern (no pivoting — `-W_inner` is
  column diagonally dominant for monotone flows);
* for every L/U entry the ordered list of `L[i,k]*U[k,j]` products it subtracts and a
  dependency level, so the numeric phase is a sequence of data-parallel levels in which each
  entry is written by exactly one thread: no atomics, bit-reproducible;
* level schedules for the forward and backward substitutions.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .interp import as_i32


@dataclass
class Symbolic:
    n: int
    perm: np.ndarray          # new index -> old index
    iperm: np.ndarray         # old index -> new index
    lu_row: np.ndarray        # permuted row of each LU entry (entries sorted by level)
    lu_col: np.ndarray
    lu_wsrc: np.ndarray       # index of the W entry feeding this LU entry, -1 for fill
    lu_pivot: np.ndarray      # for L entries: index of U[k,k]; -1 for U entries
    lu_prod_ptr: np.ndarray
    lu_prod_l: np.ndarray
    lu_prod_u: np.ndarray
    lu_level_ptr: np.ndarray
    lu_diag: np.ndarray       # index of U[i,i] per permuted row
    fwd_level_ptr: np.ndarray
    fwd_row: np.ndarray       # permuted rows sorted by level
    fwd_ptr: np.ndarray       # per position in fwd_row
    fwd_l: np.ndarray         # L entry index
    fwd_k: np.ndarray         # permuted column (index into y)
    bwd_level_ptr: np.ndarray
    bwd_row: np.ndarray
    bwd_ptr: np.ndarray
    bwd_u: np.ndarray
    bwd_k: np.ndarray
    n_rounds: int
    # "LDU" variant used by the shared-memory kernel: the same L/U slots, but L is kept unscaled
    # and the pivot reciprocals r_k = 1/U[k,k] are stored in the diagonal slots, so an entry only
    # depends on pivots of earlier rounds: one level per round, no division on the critical path.
    # slot(i,j) -= slot(i,k) * r_k * slot(k,j);  forward: yhat_i = r_i (b_i - sum L~[i,k] yhat_k);
    # backward: x_i = yhat_i - r_i sum U[i,j] x_j.
    ldu_wsrc: np.ndarray = None      # W entry feeding each slot (-1 fill); slots sorted by round
    ldu_flag: np.ndarray = None      # 1 for diagonal slots
    ldu_prod_ptr: np.ndarray = None
    ldu_prod_l: np.ndarray = None
    ldu_prod_u: np.ndarray = None
    ldu_prod_p: np.ndarray = None    # slot of r_k of the product's pivot
    ldu_level_ptr: np.ndarray = None
    ldu_diag: np.ndarray = None      # slot of r_i per permuted row
    ldu_fwd_level_ptr: np.ndarray = None
    ldu_fwd_row: np.ndarray = None
    ldu_fwd_ptr: np.ndarray = None
    ldu_fwd_l: np.ndarray = None
    ldu_fwd_k: np.ndarray = None
    ldu_bwd_level_ptr: np.ndarray = None
    ldu_bwd_row: np.ndarray = None
    ldu_bwd_ptr: np.ndarray = None
    ldu_bwd_u: np.ndarray = None
    ldu_bwd_k: np.ndarray = None

    @property
    def n_lu(self) -> int:
        return len(self.lu_row)


def rake_compress_order(n: int, adj: list[set[int]]):
    """Return (elimination order, higher-ordered neighbour set of each pivot, #rounds,
    round of each vertex).

    `adj` is consumed (it becomes the quotient graph)."""
    alive = np.ones(n, dtype=bool)
    order: list[int] = []
    struct: list[list[int]] = [[] for _ in range(n)]
    deg = np.array([len(a) for a in adj], dtype=np.int64)
    round_of = np.zeros(n, dtype=np.int64)
    rounds = 0
    remaining = n
    while remaining:
        rounds += 1
        live = np.flatnonzero(alive)
        d = deg[live]
        # candidates: degree up to 2*min + 2.  The slack lets the dense-ish core that chords
        # leave behind be eliminated in a handful of rounds (independent sets of several
        # vertices) instead of one vertex per round, for a few percent more fill.
        thr = max(2 * int(d.min()) + 2, 2)
        cand = live[d <= thr]
        cand = cand[np.argsort(deg[cand], kind="stable")]
        blocked: set[int] = set()
        picked: list[int] = []
        for v in cand.tolist():
            if v in blocked:
                continue
            picked.append(v)
            blocked.update(adj[v])
        for v in picked:
            nb = adj[v]
            struct[v] = sorted(nb)
            for a in nb:
                s = adj[a]
                s.discard(v)
                s.update(nb)
                s.discard(a)
                deg[a] = len(s)
            adj[v] = set()
            alive[v] = False
            round_of[v] = rounds - 1
            order.append(v)
        remaining -= len(picked)
    return order, struct, rounds, round_of


def analyse(n: int, wrow_ptr: np.ndarray, wcol: np.ndarray) -> Symbolic:
    """Symbolic LU of an n x n CSR pattern (diagonal assumed present)."""
    adj: list[set[int]] = [set() for _ in range(n)]
    wpos: dict[tuple[int, int], int] = {}
    for i in range(n):
        for k in range(int(wrow_ptr[i]), int(wrow_ptr[i + 1])):
            j = int(wcol[k])
            wpos[(i, j)] = k
            if i != j:
                adj[i].add(j)
                adj[j].add(i)
    order, struct_old, rounds, round_old = rake_compress_order(n, adj)
    perm = np.array(order, dtype=np.int64)
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)

    # entries: U row k = {k} + S_k, L column k = S_k (all in permuted indices)
    ent_row: list[int] = []
    ent_col: list[int] = []
    ent_idx: dict[tuple[int, int], int] = {}
    S: list[list[int]] = []
    for k in range(n):
        sk = sorted(int(iperm[v]) for v in struct_old[order[k]])
        S.append(sk)
        for j in [k] + sk:
            ent_idx[(k, j)] = len(ent_row)
            ent_row.append(k)
            ent_col.append(j)
        for i in sk:
            ent_idx[(i, k)] = len(ent_row)
            ent_row.append(i)
            ent_col.append(k)
    n_lu = len(ent_row)
    prods: list[list[tuple[int, int]]] = [[] for _ in range(n_lu)]
    level = np.zeros(n_lu, dtype=np.int64)
    acc = np.full(n_lu, -1, dtype=np.int64)
    pivot = np.full(n_lu, -1, dtype=np.int64)
    diag = np.empty(n, dtype=np.int64)
    for k in range(n):
        sk = S[k]
        dk = ent_idx[(k, k)]
        diag[k] = dk
        level[dk] = acc[dk] + 1
        for j in sk:
            e = ent_idx[(k, j)]
            level[e] = acc[e] + 1
        for i in sk:
            e = ent_idx[(i, k)]
            pivot[e] = dk
            level[e] = max(acc[e], level[dk]) + 1
        for i in sk:
            el = ent_idx[(i, k)]
            ll = level[el]
            for j in sk:
                eu = ent_idx[(k, j)]
                t = ent_idx[(i, j)]
                prods[t].append((el, eu))
                lv = max(ll, level[eu])
                if lv > acc[t]:
                    acc[t] = lv

    # sort entries by level (stable) and renumber
    order_e = np.argsort(level, kind="stable")
    new_of = np.empty(n_lu, dtype=np.int64)
    new_of[order_e] = np.arange(n_lu)
    n_levels = int(level.max()) + 1 if n_lu else 0
    level_ptr = np.searchsorted(level[order_e], np.arange(n_levels + 1), side="left")
    lu_row = np.array(ent_row, dtype=np.int64)[order_e]
    lu_col = np.array(ent_col, dtype=np.int64)[order_e]
    lu_pivot = np.where(pivot[order_e] >= 0, new_of[np.maximum(pivot[order_e], 0)], -1)
    lu_wsrc = np.full(n_lu, -1, dtype=np.int64)
    for e_new, e_old in enumerate(order_e.tolist()):
        i_old, j_old = order[ent_row[e_old]], order[ent_col[e_old]]
        lu_wsrc[e_new] = wpos.get((i_old, j_old), -1)
    prod_ptr = [0]
    prod_l: list[int] = []
    prod_u: list[int] = []
    for e_old in order_e.tolist():
        for (el, eu) in prods[e_old]:
            prod_l.append(int(new_of[el]))
            prod_u.append(int(new_of[eu]))
        prod_ptr.append(len(prod_l))
    lu_diag = new_of[diag]

    # substitution schedules
    Lrow: list[list[tuple[int, int]]] = [[] for _ in range(n)]  # (col k, entry)
    Urow: list[list[tuple[int, int]]] = [[] for _ in range(n)]
    for k in range(n):
        for i in S[k]:
            Lrow[i].append((k, int(new_of[ent_idx[(i, k)]])))
        for j in S[k]:
            Urow[k].append((j, int(new_of[ent_idx[(k, j)]])))
    lvl_y = np.zeros(n, dtype=np.int64)
    for i in range(n):
        if Lrow[i]:
            lvl_y[i] = 1 + max(lvl_y[k] for k, _ in Lrow[i])
    lvl_x = np.zeros(n, dtype=np.int64)
    for i in range(n - 1, -1, -1):
        if Urow[i]:
            lvl_x[i] = 1 + max(lvl_x[j] for j, _ in Urow[i])

    def schedule(lvl, rows):
        idx = np.argsort(lvl, kind="stable")
        nl = int(lvl.max()) + 1 if n else 0
        lptr = np.searchsorted(lvl[idx], np.arange(nl + 1), side="left")
        ptr = [0]
        ent: list[int] = []
        col: list[int] = []
        for i in idx.tolist():
            for k, e in sorted(rows[i]):
                ent.append(e)
                col.append(k)
            ptr.append(len(ent))
        return as_i32(lptr), as_i32(idx), as_i32(ptr), as_i32(ent), as_i32(col)

    f_lptr, f_row, f_ptr, f_l, f_k = schedule(lvl_y, Lrow)
    b_lptr, b_row, b_ptr, b_u, b_k = schedule(lvl_x, Urow)

    # ---- LDU variant: slots sorted by the round of their pivot min(i, j)
    rnd = np.array([round_old[order[min(r, c)]] for r, c in zip(ent_row, ent_col)], dtype=np.int64)
    order_d = np.argsort(rnd, kind="stable")
    new_d = np.empty(n_lu, dtype=np.int64)
    new_d[order_d] = np.arange(n_lu)
    d_level_ptr = np.searchsorted(rnd[order_d], np.arange(rounds + 1), side="left")
    d_wsrc = np.full(n_lu, -1, dtype=np.int64)
    d_flag = np.zeros(n_lu, dtype=np.int64)
    dp_ptr = [0]
    dp_l: list[int] = []
    dp_u: list[int] = []
    dp_p: list[int] = []
    for e_new, e_old in enumerate(order_d.tolist()):
        i_old, j_old = order[ent_row[e_old]], order[ent_col[e_old]]
        d_wsrc[e_new] = wpos.get((i_old, j_old), -1)
        d_flag[e_new] = int(ent_row[e_old] == ent_col[e_old])
        for (el, eu) in prods[e_old]:
            kpiv = ent_col[el]  # L[i,k] sits in column k
            dp_l.append(int(new_d[el]))
            dp_u.append(int(new_d[eu]))
            dp_p.append(int(new_d[diag[kpiv]]))
        dp_ptr.append(len(dp_l))
    d_diag = new_d[diag]

    def schedule_d(lvl, rows):
        idx = np.argsort(lvl, kind="stable")
        nl = int(lvl.max()) + 1 if n else 0
        lptr = np.searchsorted(lvl[idx], np.arange(nl + 1), side="left")
        ptr = [0]
        ent: list[int] = []
        col: list[int] = []
        for i in idx.tolist():
            for k, e in sorted(rows[i]):
                ent.append(int(new_d[order_e[e]]))  # rows[] hold LU-sorted slot ids
                col.append(k)
            ptr.append(len(ent))
        return as_i32(lptr), as_i32(idx), as_i32(ptr), as_i32(ent), as_i32(col)

    df_lptr, df_row, df_ptr, df_l, df_k = schedule_d(lvl_y, Lrow)
    db_lptr, db_row, db_ptr, db_u, db_k = schedule_d(lvl_x, Urow)
    return Symbolic(
        ldu_wsrc=as_i32(d_wsrc), ldu_flag=as_i32(d_flag), ldu_prod_ptr=as_i32(dp_ptr),
        ldu_prod_l=as_i32(dp_l), ldu_prod_u=as_i32(dp_u), ldu_prod_p=as_i32(dp_p),
        ldu_level_ptr=as_i32(d_level_ptr), ldu_diag=as_i32(d_diag),
        ldu_fwd_level_ptr=df_lptr, ldu_fwd_row=df_row, ldu_fwd_ptr=df_ptr, ldu_fwd_l=df_l,
        ldu_fwd_k=df_k, ldu_bwd_level_ptr=db_lptr, ldu_bwd_row=db_row, ldu_bwd_ptr=db_ptr,
        ldu_bwd_u=db_u, ldu_bwd_k=db_k,
        n=n, perm=as_i32(perm), iperm=as_i32(iperm),
        lu_row=as_i32(lu_row), lu_col=as_i32(lu_col), lu_wsrc=as_i32(lu_wsrc),
        lu_pivot=as_i32(lu_pivot), lu_prod_ptr=as_i32(prod_ptr), lu_prod_l=as_i32(prod_l),
        lu_prod_u=as_i32(prod_u), lu_level_ptr=as_i32(level_ptr), lu_diag=as_i32(lu_diag),
        fwd_level_ptr=f_lptr, fwd_row=f_row, fwd_ptr=f_ptr, fwd_l=f_l, fwd_k=f_k,
        bwd_level_ptr=b_lptr, bwd_row=b_row, bwd_ptr=b_ptr, bwd_u=b_u, bwd_k=b_k,
        n_rounds=rounds,
    )


# ------------------------------------------------------------------ reference numeric phase
def numeric_factor(sym: Symbolic, wvals: np.ndarray) -> np.ndarray:
    """NumPy restatement of the device numeric phase (used by the CPU tests of the schedule)."""
    lu = np.zeros(sym.n_lu)
    for lv in range(len(sym.lu_level_ptr) - 1):
        for e in range(sym.lu_level_ptr[lv], sym.lu_level_ptr[lv + 1]):
            v = wvals[sym.lu_wsrc[e]] if sym.lu_wsrc[e] >= 0 else 0.0
            for q in range(sym.lu_prod_ptr[e], sym.lu_prod_ptr[e + 1]):
                v -= lu[sym.lu_prod_l[q]] * lu[sym.lu_prod_u[q]]
            if sym.lu_pivot[e] >= 0:
                v /= lu[sym.lu_pivot[e]]
            lu[e] = v
    return lu


def numeric_solve(sym: Symbolic, lu: np.ndarray, b: np.ndarray) -> np.ndarray:
    x = b[sym.perm].astype(np.float64).copy()
    for lv in range(len(sym.fwd_level_ptr) - 1):
        for pos in range(sym.fwd_level_ptr[lv], sym.fwd_level_ptr[lv + 1]):
            i = sym.fwd_row[pos]
            v = x[i]
            for q in range(sym.fwd_ptr[pos], sym.fwd_ptr[pos + 1]):
                v -= lu[sym.fwd_l[q]] * x[sym.fwd_k[q]]
            x[i] = v
    for lv in range(len(sym.bwd_level_ptr) - 1):
        for pos in range(sym.bwd_level_ptr[lv], sym.bwd_level_ptr[lv + 1]):
            i = sym.bwd_row[pos]
            v = x[i]
            for q in range(sym.bwd_ptr[pos], sym.bwd_ptr[pos + 1]):
                v -= lu[sym.bwd_u[q]] * x[sym.bwd_k[q]]
            x[i] = v / lu[sym.lu_diag[i]]
    out = np.empty_like(x)
    out[sym.perm] = x
    return out


def numeric_factor_ldu(sym: Symbolic, wvals: np.ndarray) -> np.ndarray:
    """NumPy restatement of the LDU numeric phase of the shared-memory kernel."""
    lu = np.zeros(sym.n_lu)
    for lv in range(len(sym.ldu_level_ptr) - 1):
        lo, hi = sym.ldu_level_ptr[lv], sym.ldu_level_ptr[lv + 1]
        new = {}
        for e in range(lo, hi):
            v = wvals[sym.ldu_wsrc[e]] if sym.ldu_wsrc[e] >= 0 else 0.0
            for q in range(sym.ldu_prod_ptr[e], sym.ldu_prod_ptr[e + 1]):
                v -= (lu[sym.ldu_prod_l[q]] * lu[sym.ldu_prod_p[q]]) * lu[sym.ldu_prod_u[q]]
            new[e] = 1.0 / v if sym.ldu_flag[e] else v
        for e, v in new.items():  # a level only reads earlier levels
            lu[e] = v
    return lu


def numeric_solve_ldu(sym: Symbolic, lu: np.ndarray, b: np.ndarray) -> np.ndarray:
    x = b[sym.perm].astype(np.float64).copy()
    for lv in range(len(sym.ldu_fwd_level_ptr) - 1):
        for pos in range(sym.ldu_fwd_level_ptr[lv], sym.ldu_fwd_level_ptr[lv + 1]):
            i = sym.ldu_fwd_row[pos]
            v = x[i]
            for q in range(sym.ldu_fwd_ptr[pos], sym.ldu_fwd_ptr[pos + 1]):
                v -= lu[sym.ldu_fwd_l[q]] * x[sym.ldu_fwd_k[q]]
            x[i] = v * lu[sym.ldu_diag[i]]
    for lv in range(1, len(sym.ldu_bwd_level_ptr) - 1):
        for pos in range(sym.ldu_bwd_level_ptr[lv], sym.ldu_bwd_level_ptr[lv + 1]):
            i = sym.ldu_bwd_row[pos]
            v = 0.0
            for q in range(sym.ldu_bwd_ptr[pos], sym.ldu_bwd_ptr[pos + 1]):
                v += lu[sym.ldu_bwd_u[q]] * x[sym.ldu_bwd_k[q]]
            x[i] = x[i] - lu[sym.ldu_diag[i]] * v
    out = np.empty_like(x)
    out[sym.perm] = x
    return out


# ------------------------------------------------------------------ entry schedule (tile kernels)
# The tile kernels (`rbs_ent_levels` in csrc/rbs_kernels.cuh) keep one member tile's factor slots
# in shared memory as S[slot][8 members] and run the whole solve as a sequence of *levels*
# separated by block barriers.  Four threads own one entry of the level (each thread two of the
# tile's eight members, one 128-bit shared-memory access per operand) and accumulate its products:
#
#     acc = sum_q (S[l_q] * S[p_q]) * S[u_q];   v = S[e] (* S[r] if SCALE) - acc;   S[e] = 1/v if RCP else v
#
# In levels with few entries and long product lists a *team* of 2^team_log entry lanes shares one
# entry (products strided over the team, partial sums combined by a fixed shuffle butterfly).
# Slots: [0, n_lu) the LDU factor slots; [n_lu, n_lu + n) the right-hand side column y (permuted
# rows), which is eliminated together with the matrix (forward substitution fused into the
# factorisation) and holds the solution after the backward levels.
#   elimination level r : slot(i,j) -= (slot(i,k) r_k) slot(k,j), diagonal -> reciprocal,
#                         y_i -= (slot(i,k) r_k) y_k
#   backward level      : x_i = y_i r_i - sum_j (slot(i,j) r_i) x_j
ENT_RCP = 1 << 8
ENT_SCALE = 1 << 9
ENT_FRESH = 1 << 10  # first write of a fill-in that took over a dead slot: the old content is not read
ENT_LANES = 256  # entry lanes per block (1024 threads, four per entry: two members each)


@dataclass
class EntrySched:
    ent_e: np.ndarray      # target slot per entry, entries sorted by level
    ent_q0: np.ndarray     # first product (or, for a SCALE entry without products, the r slot)
    ent_cf: np.ndarray     # product count (low 8 bits) | ENT_RCP | ENT_SCALE
    prod_l: np.ndarray
    prod_p: np.ndarray
    prod_u: np.ndarray
    lvl_ptr: np.ndarray    # entries of level lv: [lvl_ptr[lv], lvl_ptr[lv + 1])
    team_log: np.ndarray   # per level: log2 of the lanes that share one entry
    n_elim: int
    n_bwd: int
    n_slots: int
    y0: int
    # compact form (slots of dead entries taken over by later fill-ins): the original entries of W
    # to load, `load_e[k]` = LU entry, `load_s[k]` = its slot; None = every LU entry e in slot e
    load_e: np.ndarray | None = None
    load_s: np.ndarray | None = None

    def smem_bytes(self, tile: int = 8) -> int:
        idx = 2 * (self.ent_e.size * 3 + self.prod_l.size * 3 + self.lvl_ptr.size + self.team_log.size)
        return self.n_slots * tile * 8 + idx + 64


def _compact_slots(sym: Symbolic, all_levels: list[list], n_ids: int):
    """Slot assignment by liveness: an LU entry's slot is free again after the last level that
    reads or writes it (the multipliers of a column die with their elimination round), and a
    fill-in is only born in the level of its first update.  Every level ends with a block barrier,
    so a fill-in may take over any slot whose last use lies in an earlier level.  Greedy in birth
    order is optimal for intervals: the slot count is the peak number of live entries.
    Returns (slot of every id, number of LU slots, ids whose first write is FRESH)."""
    n, n_lu = sym.n, sym.n_lu
    first_w = np.full(n_ids, 1 << 30, dtype=np.int64)
    last_use = np.full(n_ids, -1, dtype=np.int64)
    for lv, lvl in enumerate(all_levels):
        for (e, prods, flags, r) in lvl:
            first_w[e] = min(first_w[e], lv)
            last_use[e] = max(last_use[e], lv)
            for trip in prods:
                for s_ in trip:
                    last_use[s_] = max(last_use[s_], lv)
            if flags & ENT_SCALE:
                rr = prods[0][1] if prods else r
                last_use[rr] = max(last_use[rr], lv)
    # loaded from the assembled values: the entries of W, and fill-ins of the symbolic pattern
    # that no product ever updates (they stay zero and may still be read)
    orig = (np.asarray(sym.ldu_wsrc)[:n_lu] >= 0) | (first_w[:n_lu] >= (1 << 30))
    slot_of = np.full(n_ids, -1, dtype=np.int64)
    nxt = 0
    pool: list[tuple[int, int]] = []  # (last use, slot) of the ids placed so far
    for e in range(n_lu):
        if orig[e]:
            slot_of[e] = nxt
            pool.append((int(last_use[e]), nxt))
            nxt += 1
    import heapq

    heapq.heapify(pool)
    fresh = set()
    fills = [e for e in range(n_lu) if not orig[e]]
    for e in sorted(fills, key=lambda e: (first_w[e], e)):
        b = int(first_w[e])
        if pool and pool[0][0] < b:
            _, sl = heapq.heappop(pool)
        else:
            sl = nxt
            nxt += 1
        slot_of[e] = sl
        fresh.add(e)
        heapq.heappush(pool, (int(last_use[e]), sl))
    for i in range(n):
        slot_of[n_lu + i] = nxt + i
    return slot_of, nxt, fresh, [e for e in range(n_lu) if orig[e]]


def build_entries(sym: Symbolic, compact: bool = False) -> EntrySched | None:
    """Entry schedule of the LDU elimination (+ fused forward substitution) and the backward
    substitution.  None when an index does not fit 16 bits.  `compact`: fill-ins take over the
    slots of dead entries (`_compact_slots`); the factors are then not available afterwards, so
    this form serves the full-Newton path only (every iteration refactors)."""
    n, n_lu = sym.n, sym.n_lu
    y0 = n_lu
    n_slots = n_lu + n
    diag = sym.ldu_diag
    n_elim = len(sym.ldu_level_ptr) - 1
    levels: list[list] = [[] for _ in range(n_elim)]
    lvl_of_slot = np.empty(max(n_lu, 1), dtype=np.int64)
    for lv in range(n_elim):
        lo, hi = int(sym.ldu_level_ptr[lv]), int(sym.ldu_level_ptr[lv + 1])
        lvl_of_slot[lo:hi] = lv
        for e in range(lo, hi):
            q0, q1 = int(sym.ldu_prod_ptr[e]), int(sym.ldu_prod_ptr[e + 1])
            is_d = bool(sym.ldu_flag[e])
            if q1 == q0 and not is_d:
                continue  # plain copy of W: already in place
            prods = [(int(sym.ldu_prod_l[q]), int(sym.ldu_prod_p[q]), int(sym.ldu_prod_u[q]))
                     for q in range(q0, q1)]
            levels[lv].append((e, prods, ENT_RCP if is_d else 0, 0))
    for pos in range(n):
        i = int(sym.ldu_fwd_row[pos])
        q0, q1 = int(sym.ldu_fwd_ptr[pos]), int(sym.ldu_fwd_ptr[pos + 1])
        if q1 == q0:
            continue
        prods = [(int(sym.ldu_fwd_l[q]), int(diag[int(sym.ldu_fwd_k[q])]), y0 + int(sym.ldu_fwd_k[q]))
                 for q in range(q0, q1)]
        levels[int(lvl_of_slot[diag[i]])].append((y0 + i, prods, 0, 0))
    n_bwd = len(sym.ldu_bwd_level_ptr) - 1
    blevels: list[list] = [[] for _ in range(n_bwd)]
    for lv in range(n_bwd):
        for pos in range(int(sym.ldu_bwd_level_ptr[lv]), int(sym.ldu_bwd_level_ptr[lv + 1])):
            i = int(sym.ldu_bwd_row[pos])
            q0, q1 = int(sym.ldu_bwd_ptr[pos]), int(sym.ldu_bwd_ptr[pos + 1])
            r_i = int(diag[i])
            prods = [(int(sym.ldu_bwd_u[q]), r_i, y0 + int(sym.ldu_bwd_k[q])) for q in range(q0, q1)]
            blevels[lv].append((y0 + i, prods, ENT_SCALE, r_i))
    load_e = load_s = None
    if compact:
        slot_of, y0, fresh, loaded = _compact_slots(sym, levels + blevels, n_lu + n)
        n_slots = y0 + n
        seen: set[int] = set()

        def remap(lvl):
            out = []
            for (e, prods, flags, r) in lvl:
                if e in fresh and e not in seen:
                    seen.add(e)
                    flags |= ENT_FRESH
                out.append((int(slot_of[e]), [tuple(int(slot_of[x]) for x in t) for t in prods], flags,
                            int(slot_of[r]) if flags & ENT_SCALE else r))
            return out

        levels = [remap(l) for l in levels]
        blevels = [remap(l) for l in blevels]
        load_e, load_s = as_i32(loaded), as_i32([slot_of[e] for e in loaded])
    ent_e, ent_q0, ent_cf, pl, pp, pu, lvl_ptr, team = [], [], [], [], [], [], [0], []
    def assign(lvl, tl):
        """Lane k of the block takes entries k, k + L, k + 2L, ... of the level (L lanes, or teams
        of 2^tl lanes).  Pass by pass the biggest entry goes to the least loaded lane; lanes are
        then ordered by entry count so that every pass is a prefix.  Returns the entry order and
        the critical lane's cost (instructions)."""
        L = ENT_LANES >> tl

        def cost(t):
            return 10 + 14 * tl + 18 * ((len(t[1]) + (1 << tl) - 1) >> tl) + (50 if t[2] & ENT_RCP else 0)

        lanes: list[list] = [[] for _ in range(min(L, max(len(lvl), 1)))]
        load = [0] * len(lanes)
        todo = sorted(lvl, key=lambda t: -cost(t))
        for p0 in range(0, len(todo), len(lanes)):
            batch = todo[p0:p0 + len(lanes)]
            for ent, w in zip(batch, sorted(range(len(lanes)), key=lambda k: load[k])):
                lanes[w].append(ent)
                load[w] += cost(ent)
        order = sorted(range(len(lanes)), key=lambda k: (-len(lanes[k]), -load[k]))
        lanes = [lanes[k] for k in order]
        ordered = []
        for ps in range(max((len(x) for x in lanes), default=0)):
            ordered.extend(x[ps] for x in lanes if len(x) > ps)
        return ordered, max(load, default=0)

    for lvl in levels + blevels:
        best = min(range(4), key=lambda tl: (assign(lvl, tl)[1], tl))
        team.append(best)
        ordered, _ = assign(lvl, best)
        for (e, prods, flags, r) in ordered:
            if len(prods) > 255:
                return None
            ent_e.append(e)
            ent_q0.append(len(pl) if prods else r)
            ent_cf.append(len(prods) | flags)
            for (l, p, u) in prods:
                pl.append(l); pp.append(p); pu.append(u)
        lvl_ptr.append(len(ent_e))
    if max(n_slots, len(pl), len(ent_e)) >= 65536:
        return None
    return EntrySched(ent_e=as_i32(ent_e), ent_q0=as_i32(ent_q0), ent_cf=as_i32(ent_cf),
                      prod_l=as_i32(pl), prod_p=as_i32(pp), prod_u=as_i32(pu),
                      lvl_ptr=as_i32(lvl_ptr), team_log=as_i32(team), n_elim=n_elim, n_bwd=n_bwd,
                      n_slots=n_slots, y0=y0, load_e=load_e, load_s=load_s)


def numeric_entries(sym: Symbolic, E: EntrySched, wvals: np.ndarray, b: np.ndarray) -> np.ndarray:
    """NumPy emulation of `rbs_ent_levels` for one member (same summation order)."""
    S = np.zeros(E.n_slots)
    if E.load_e is None:
        for e in range(sym.n_lu):
            S[e] = wvals[sym.ldu_wsrc[e]] if sym.ldu_wsrc[e] >= 0 else 0.0
    else:
        S[:] = np.nan  # what a slot held before is never read
        for e, sl in zip(E.load_e, E.load_s):
            S[sl] = wvals[sym.ldu_wsrc[e]] if sym.ldu_wsrc[e] >= 0 else 0.0
    S[E.y0:E.y0 + sym.n] = b[sym.perm]
    for lv in range(E.n_elim + E.n_bwd):
        ts = 1 << int(E.team_log[lv])
        new = {}
        for k in range(int(E.lvl_ptr[lv]), int(E.lvl_ptr[lv + 1])):
            e, q0, cf = int(E.ent_e[k]), int(E.ent_q0[k]), int(E.ent_cf[k])
            cnt = cf & 0xFF
            part = []
            for t in range(ts):
                acc = 0.0
                for q in range(q0 + t, q0 + cnt, ts) if cnt else ():
                    acc += (S[E.prod_l[q]] * S[E.prod_p[q]]) * S[E.prod_u[q]]
                part.append(acc)
            step = 1
            while step < ts:  # butterfly: lane t adds lane t ^ step
                part = [part[t] + part[t ^ step] for t in range(ts)]
                step <<= 1
            old = 0.0 if cf & ENT_FRESH else S[e]
            if cf & ENT_SCALE:
                old = old * S[E.prod_p[q0] if cnt else q0]
            v = old - part[0]
            new[e] = 1.0 / v if cf & ENT_RCP else v
        for e, v in new.items():  # a level only reads earlier levels
            S[e] = v
    out = np.empty(sym.n)
    out[sym.perm] = S[E.y0:E.y0 + sym.n]
    return out
