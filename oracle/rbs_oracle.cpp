// rbs_oracle.cpp — CPU restatement of Ribasim's implicit-integration inner loop.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under ribasim_b200/ may include, link or call
// this file.  It is used by tests/, by __graft_entry__.smoke() and by bench.py's
// cpu_baseline / --impl reference legs as the checker and the CPU baseline.
//
// The reference (Julia + SciML packages) cannot be compiled in this container, so this is a
// scalar fp64 restatement that follows the reference function by function (file:line cited
// at each function, relative to /root/reference).  The RHS is templated on the number type
// and instantiated with `double` and with a one-partial dual number, which is how the
// reference obtains its Jacobian (ForwardDiff, chunk size 1, core/src/config.jl:394-401).
//
// Pinning: reduction_factor, relaxed_root, the profile inverse, the state layout, the
// incidence matrix A, the Jacobian sparsity pattern and end-of-run storages are checked
// against the reference's golden values (tests/test_oracle_golden.py), as are the reference's
// own model tests (equations, control records, run_models, time) and the five PCHIP values it
// pins (validation_test.jl:7-15).  Per-entry RHS / Jacobian values and the interior PCHIP slopes
// of longer tables are NOT pinned by any reference test ("parity unpinned" for those; see
// DESIGN.md).
//
// Build: oracle/Makefile -> oracle/librbs_oracle.so   (compiled with -ffp-contract=off so
// that no FMA contraction changes the reference's operation order).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// number types
// ------------------------------------------------------------------------------------------
struct Dual {
    double v, d;
    Dual() : v(0), d(0) {}
    Dual(double v_) : v(v_), d(0) {}
    Dual(double v_, double d_) : v(v_), d(d_) {}
};
inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
inline Dual operator-(Dual a) { return {-a.v, -a.d}; }
inline Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
inline Dual operator/(Dual a, Dual b) {
    double q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
inline Dual& operator+=(Dual& a, Dual b) { a = a + b; return a; }
inline Dual& operator-=(Dual& a, Dual b) { a = a - b; return a; }
inline Dual& operator*=(Dual& a, Dual b) { a = a * b; return a; }
inline bool operator<(Dual a, Dual b) { return a.v < b.v; }
inline bool operator>(Dual a, Dual b) { return a.v > b.v; }
inline double val(double x) { return x; }
inline double val(Dual x) { return x.v; }
inline double xsqrt(double x) { return std::sqrt(x); }
inline Dual xsqrt(Dual x) { double s = std::sqrt(x.v); return {s, x.d / (2.0 * s)}; }
inline double xcbrt(double x) { return std::cbrt(x); }
inline Dual xcbrt(Dual x) { double c = std::cbrt(x.v); return {c, x.d / (3.0 * (c * c))}; }
inline double xabs(double x) { return std::fabs(x); }
inline Dual xabs(Dual x) { return x.v < 0 ? -x : x; }

template <class T> inline T jl_max(T a, double b) {  // Julia max propagates NaN
    if (val(a) != val(a)) return a;
    return val(a) > b ? a : T(b);
}
template <class T> inline T jl_min(T a, T b) {
    if (val(a) != val(a)) return a;
    if (val(b) != val(b)) return b;
    return val(b) < val(a) ? b : a;
}
template <class T> inline T jl_clamp(T x, T lo, T hi) {  // Base.clamp: x > hi ? hi : (x < lo ? lo : x)
    if (val(x) > val(hi)) return hi;
    if (val(x) < val(lo)) return lo;
    return x;
}

// core/src/util.jl:312-321
template <class T> inline T reduction_factor(T x, double threshold) {
    if (val(x) < 0) return T(0.0);
    if (val(x) < threshold) {
        T xs = x / T(threshold);
        return (T(-2.0) * xs + T(3.0)) * (xs * xs);
    }
    return T(1.0);
}

// core/src/util.jl:679-685
template <class T> inline T relaxed_root(T x, T threshold) {
    if (val(xabs(x)) < val(threshold)) {
        T r = x / threshold;
        return T(0.25) * (x / xsqrt(threshold)) * (T(5.0) - r * r);
    }
    T s = xsqrt(xabs(x));
    return val(x) > 0 ? s : (val(x) < 0 ? -s : T(0.0) * s);
}

enum NodeType { BASIN = 0, LEVEL_BOUNDARY = 1, TERMINAL = 2, TRC = 3, PUMP = 4, OUTLET = 5,
                USER_DEMAND = 6, LR = 7, MANNING = 8, FLOW_BOUNDARY = 9, OTHER = 10 };
enum { CONTROL_NONE = 0, CONTROL_CONTINUOUS = 1, CONTROL_PID = 2 };
enum { SV_BASIN_LEVEL = 0, SV_BASIN_STORAGE = 1, SV_DU_STATE = 2, SV_CONSTANT = 3 };

struct Store {
    std::map<std::string, std::vector<double>> d;
    std::map<std::string, std::vector<int32_t>> i;
};

// PCHIP table: restatement of DataInterpolations' PCHIPInterpolation / CubicHermiteSpline
// (third-party, Manifest.toml:375-381; call sites core/src/util.jl:134-140, read.jl:1188-1193)
struct Pchip {
    std::vector<double> t, u, du, c1, c2;
    bool left_linear = false, right_linear = false;
    double slope_left = 0, slope_right = 0;
    void build() {
        size_t n = t.size();
        // Fritsch-Carlson / Moler pchip slopes; end rule pinned by
        // core/test/validation_test.jl:7-15 (itp(1.5) = 0.03125, itp(3.0) = 0.25)
        std::vector<double> h(n - 1), delta(n - 1);
        for (size_t k = 0; k + 1 < n; k++) {
            h[k] = t[k + 1] - t[k];
            delta[k] = (u[k + 1] - u[k]) / h[k];
        }
        du.assign(n, 0.0);
        auto sgn = [](double x) { return (x > 0) - (x < 0); };
        auto pchip_end = [&](double h1, double h2, double del1, double del2) {
            double d = ((2 * h1 + h2) * del1 - h1 * del2) / (h1 + h2);
            if (sgn(d) != sgn(del1)) return 0.0;
            if (sgn(del1) != sgn(del2) && std::fabs(d) > std::fabs(3 * del1)) return 3 * del1;
            return d;
        };
        if (n == 2) { du[0] = du[1] = delta[0]; }
        else {
            for (size_t k = 1; k + 1 < n; k++) {
                double a = delta[k - 1], b = delta[k];
                if (a == 0.0 || b == 0.0 || (a > 0) != (b > 0)) du[k] = 0.0;
                else {
                    double w1 = 2 * h[k] + h[k - 1], w2 = h[k] + 2 * h[k - 1];
                    du[k] = a * b * (w1 + w2) / (w1 * b + w2 * a);
                }
            }
            du[0] = pchip_end(h[0], h[1], delta[0], delta[1]);
            du[n - 1] = pchip_end(h[n - 2], h[n - 3], delta[n - 2], delta[n - 3]);
        }
        c1.resize(n - 1); c2.resize(n - 1);
        for (size_t i = 0; i + 1 < n; i++) {
            double dt = t[i + 1] - t[i];
            c1[i] = (u[i + 1] - u[i] - du[i] * dt) / (dt * dt);
            c2[i] = (du[i + 1] - du[i] - 2 * c1[i] * dt) / (dt * dt);
        }
        slope_left = left_linear ? du[0] : 0.0;
        if (right_linear) {
            size_t i = n - 2;
            double d0 = t[n - 1] - t[i];
            slope_right = du[i] + d0 * (d0 * c2[i] + 2 * (c1[i] + 0.0 * c2[i]));
        } else slope_right = 0.0;
    }
    template <class T> T eval(T x) const {
        size_t n = t.size();
        double xv = val(x);
        if (xv < t[0]) return T(u[0]) + T(slope_left) * (x - T(t[0]));
        if (xv > t[n - 1]) return T(u[n - 1]) + T(slope_right) * (x - T(t[n - 1]));
        size_t idx = std::upper_bound(t.begin(), t.end(), xv) - t.begin();  // searchsortedlast
        if (idx < 1) idx = 1;
        if (idx > n - 1) idx = n - 1;
        size_t i = idx - 1;
        T d0 = x - T(t[i]);
        T d1 = x - T(t[i + 1]);
        T out = T(u[i]) + d0 * T(du[i]);
        out += (d0 * d0) * (T(c1[i]) + d1 * T(c2[i]));
        return out;
    }
};

// banded/sparse helper for the CPU baseline: LU without pivoting on a CSR-like dense-row
// skyline is overkill here; the baseline uses the same dense solver for small n_r and a
// sparse Gaussian elimination with a static minimum-degree order for large n_r.
struct SparseLU {
    int n = 0;
    std::vector<int> order;                      // elimination order
    std::vector<std::vector<int>> cols;          // per row: sorted column indices (filled pattern)
    std::vector<std::vector<double>> vals;
    std::vector<int> pos_of;                     // position in order
    void analyse(int n_, const std::vector<std::vector<int>>& pattern) {
        n = n_;
        std::vector<std::vector<int>> adj(n);
        for (int i = 0; i < n; i++)
            for (int j : pattern[i]) if (i != j) { adj[i].push_back(j); adj[j].push_back(i); }
        for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
        std::vector<char> done(n, 0);
        order.clear(); pos_of.assign(n, 0);
        std::vector<std::vector<int>> fill = adj;
        // simple minimum-degree with lazy degree buckets
        std::multimap<int, int> heap;
        std::vector<int> deg(n);
        for (int i = 0; i < n; i++) { deg[i] = (int)fill[i].size(); heap.insert({deg[i], i}); }
        while (!heap.empty()) {
            auto it = heap.begin(); int d = it->first, v = it->second; heap.erase(it);
            if (done[v] || d != deg[v]) continue;
            done[v] = 1; pos_of[v] = (int)order.size(); order.push_back(v);
            std::vector<int> nb;
            for (int a : fill[v]) if (!done[a]) nb.push_back(a);
            for (int a : nb) {
                auto& fa = fill[a];
                std::vector<int> merged;
                std::set_union(fa.begin(), fa.end(), nb.begin(), nb.end(), std::back_inserter(merged));
                merged.erase(std::remove_if(merged.begin(), merged.end(),
                                            [&](int x) { return x == a || done[x]; }), merged.end());
                fa.swap(merged);
                deg[a] = (int)fa.size();
                heap.insert({deg[a], a});
            }
            fill[v] = nb;  // higher-ordered neighbours of v
        }
        cols.assign(n, {});
        for (int v = 0; v < n; v++) {
            // row pattern in permuted space: all k with pos < pos_of[v] that have v as a higher
            // neighbour, plus v's own higher neighbours, plus the diagonal
            cols[pos_of[v]].push_back(pos_of[v]);
            for (int a : fill[v]) { cols[pos_of[v]].push_back(pos_of[a]); cols[pos_of[a]].push_back(pos_of[v]); }
        }
        for (auto& c : cols) { std::sort(c.begin(), c.end()); c.erase(std::unique(c.begin(), c.end()), c.end()); }
        vals.assign(n, {});
        for (int i = 0; i < n; i++) vals[i].assign(cols[i].size(), 0.0);
    }
    double& at(int i, int j) {  // permuted indices, must be in the pattern
        auto& c = cols[i];
        return vals[i][std::lower_bound(c.begin(), c.end(), j) - c.begin()];
    }
    void clear() { for (auto& v : vals) std::fill(v.begin(), v.end(), 0.0); }
    bool factor() {
        for (int k = 0; k < n; k++) {
            double d = at(k, k);
            if (d == 0.0 || d != d) return false;
            auto& ck = cols[k];
            size_t kk = std::lower_bound(ck.begin(), ck.end(), k) - ck.begin();
            for (size_t q = kk + 1; q < ck.size(); q++) {
                int i = ck[q];  // symmetric pattern: row i has column k
                double l = at(i, k) / d;
                at(i, k) = l;
                if (l == 0.0) continue;
                for (size_t r = kk + 1; r < ck.size(); r++) at(i, ck[r]) -= l * vals[k][r];
            }
        }
        return true;
    }
    void solve(double* b /* original order, overwritten */) {
        std::vector<double> x(n);
        for (int i = 0; i < n; i++) x[pos_of[i]] = b[i];
        for (int i = 0; i < n; i++)
            for (size_t q = 0; q < cols[i].size() && cols[i][q] < i; q++) x[i] -= vals[i][q] * x[cols[i][q]];
        for (int i = n - 1; i >= 0; i--) {
            double v = x[i];
            size_t q = std::lower_bound(cols[i].begin(), cols[i].end(), i) - cols[i].begin();
            double d = vals[i][q];
            for (q = q + 1; q < cols[i].size(); q++) v -= vals[i][q] * x[cols[i][q]];
            x[i] = v / d;
        }
        for (int i = 0; i < n; i++) b[i] = x[pos_of[i]];
    }
};

struct Model {
    Store s;
    SparseLU lu; bool lu_ready = false;
    bool finalized = false;
    int n_basin = 0, n_trc = 0, n_pump = 0, n_outlet = 0, n_ud = 0, n_ud_link = 0, n_lr = 0,
        n_manning = 0, n_pid = 0, n_cc = 0, n_lb = 0, n_fb = 0, n_prio = 0;
    int off_trc, off_pump, off_outlet, off_ud_in, off_ud_out, off_lr, off_manning, off_evap,
        off_infil, off_integral, n_state, n_reduced;
    double level_difference_threshold = 0.02;
    double tprev = 0.0;
    double h_override = -1.0;  // >= 0: exact forcing integrates over this sub-step length
    std::vector<double> zlast; double hlast = 0.0;  // last accepted increment (predictor)
    // DiscreteControl state (core/src/callback.jl:608-696)
    int n_dc = 0;
    std::vector<int> dc_truth, dc_state;
    int dc_changes = 0, dc_undefined = 0;
    std::vector<int> dc_record;  // (node index, truth bits, new control state) per change
    int merged_exact = 0;  // 1: sum the exact forcing terms first (device association)
    std::vector<Pchip> tabs;
    // colouring of J_intermediate columns
    std::vector<int> color_of_col; int n_colors = 0;
    std::vector<int> pat_rows, pat_cols;  // CSC pattern given by the caller
    // limiter memory (basin.storage_prev / level_prev: frozen when concentration is off)
    std::vector<double> storage_prev, level_prev;

    // bound views of the named arrays (rebound by bind() after every set/clone)
#define RBS_D_FIELDS(X) X(dc_sv_weight) X(dc_thr_hi) X(dc_thr_lo) X(dc_sv_const) X(pump_cs_flow_rate) X(pump_cs_min_flow_rate) X(pump_cs_max_flow_rate) X(pump_cs_min_upstream_level) X(pump_cs_max_downstream_level) X(outlet_cs_flow_rate) X(outlet_cs_min_flow_rate) X(outlet_cs_max_flow_rate) X(outlet_cs_min_upstream_level) X(outlet_cs_max_downstream_level) X(lr_cs_resistance) X(lr_cs_max_flow_rate) X(mn_cs_length) X(mn_cs_manning_n) X(mn_cs_profile_width) X(mn_cs_profile_slope) X(pid_cs_target) X(pid_cs_proportional) X(pid_cs_integral) X(pid_cs_derivative) X(cc_sv_const) X(cc_sv_weight) X(cumulative_drainage) X(cumulative_precipitation) X(cumulative_surface_runoff) X(drainage) X(fb_cumulative_flow) X(fb_incr) X(fb_rate_now) X(fb_rate_step) X(fixed_area) X(infiltration) X(lb_level) X(low_storage_threshold) X(lr_max_flow_rate) X(lr_resistance) X(mn_downstream_bottom) X(mn_length) X(mn_manning_n) X(mn_profile_slope) X(mn_profile_width) X(mn_upstream_bottom) X(pid_derivative) X(pid_integral) X(pid_proportional) X(pid_target) X(potential_evaporation) X(precipitation) X(prof_area) X(prof_cumS) X(prof_dSdh) X(prof_level) X(storage0) X(surface_runoff) X(tab_t) X(tab_u) X(trc_max_downstream_level) X(ud_allocated) X(ud_demand) X(ud_demand_static) X(ud_link_alloc) X(ud_min_level) X(ud_return_factor) X(pump_flow_rate) X(pump_min_flow_rate) X(pump_max_flow_rate) X(pump_min_upstream_level) X(pump_max_downstream_level) X(pump_flow_demand) X(outlet_flow_rate) X(outlet_min_flow_rate) X(outlet_max_flow_rate) X(outlet_min_upstream_level) X(outlet_max_downstream_level) X(outlet_flow_demand)
#define RBS_I_FIELDS(X) X(dc_cond_ptr) X(dc_csv_ptr) X(dc_sv_kind) X(dc_sv_idx) X(dc_logic_ptr) X(dc_logic) X(pump_dc) X(pump_cs_ptr) X(outlet_dc) X(outlet_cs_ptr) X(lr_dc) X(lr_cs_ptr) X(mn_dc) X(mn_cs_ptr) X(pid_dc) X(pid_cs_ptr) X(trc_dc) X(trc_cs_ptr) X(trc_cs_table) X(bin_idx) X(bin_ptr) X(bin_type) X(bout_idx) X(bout_ptr) X(bout_type) X(cc_sv_idx) X(cc_sv_kind) X(cc_sv_ptr) X(cc_table_idx) X(cc_target_idx) X(cc_target_is_outlet) X(fb_dst_idx) X(fb_dst_type) X(lr_dst_idx) X(lr_dst_type) X(lr_src_idx) X(lr_src_type) X(mn_dst_idx) X(mn_dst_type) X(mn_src_idx) X(mn_src_type) X(pid_listen_basin) X(pid_target_idx) X(pid_target_is_outlet) X(prof_ptr) X(tab_left_linear) X(tab_ptr) X(tab_right_linear) X(trc_dst_idx) X(trc_dst_type) X(trc_src_idx) X(trc_src_type) X(trc_table_idx) X(ud_demand_from_timeseries) X(ud_dst_idx) X(ud_dst_type) X(ud_has_priority) X(ud_link_ptr) X(ud_link_src_idx) X(ud_link_src_type) X(pump_control_type) X(pump_src_type) X(pump_src_idx) X(pump_dst_type) X(pump_dst_idx) X(outlet_control_type) X(outlet_src_type) X(outlet_src_idx) X(outlet_dst_type) X(outlet_dst_idx)
#define X(name) double* name = nullptr;
    RBS_D_FIELDS(X)
#undef X
#define X(name) const int32_t* name = nullptr;
    RBS_I_FIELDS(X)
#undef X
    void bind() {
#define X(name) { auto it = s.d.find(#name); name = it == s.d.end() || it->second.empty() ? nullptr : it->second.data(); }
        RBS_D_FIELDS(X)
#undef X
#define X(name) { auto it = s.i.find(#name); name = it == s.i.end() || it->second.empty() ? nullptr : it->second.data(); }
        RBS_I_FIELDS(X)
#undef X
    }
    Model() = default;
    Model(const Model& o) { *this = o; }
    Model& operator=(const Model& o) {
        s = o.s; finalized = o.finalized;
        n_basin = o.n_basin; n_trc = o.n_trc; n_pump = o.n_pump; n_outlet = o.n_outlet; n_ud = o.n_ud;
        n_ud_link = o.n_ud_link; n_lr = o.n_lr; n_manning = o.n_manning; n_pid = o.n_pid; n_cc = o.n_cc;
        n_lb = o.n_lb; n_fb = o.n_fb; n_prio = o.n_prio;
        off_trc = o.off_trc; off_pump = o.off_pump; off_outlet = o.off_outlet; off_ud_in = o.off_ud_in;
        off_ud_out = o.off_ud_out; off_lr = o.off_lr; off_manning = o.off_manning; off_evap = o.off_evap;
        off_infil = o.off_infil; off_integral = o.off_integral; n_state = o.n_state; n_reduced = o.n_reduced;
        level_difference_threshold = o.level_difference_threshold; tprev = o.tprev; h_override = o.h_override; zlast = o.zlast; hlast = o.hlast; n_dc = o.n_dc; dc_truth = o.dc_truth; dc_state = o.dc_state;
        dc_changes = o.dc_changes; dc_undefined = o.dc_undefined; dc_record = o.dc_record;
        merged_exact = o.merged_exact; tabs = o.tabs; color_of_col = o.color_of_col; n_colors = o.n_colors;
        pat_rows = o.pat_rows; pat_cols = o.pat_cols; storage_prev = o.storage_prev; level_prev = o.level_prev;
        lu = o.lu; lu_ready = o.lu_ready;
        bind();
        return *this;
    }

    const std::vector<double>& D(const char* k) const {
        auto it = s.d.find(k);
        if (it == s.d.end()) { static std::vector<double> e; return e; }
        return it->second;
    }
    std::vector<double>& Dm(const char* k) { return s.d[k]; }
    const std::vector<int32_t>& I(const char* k) const {
        auto it = s.i.find(k);
        if (it == s.i.end()) { static std::vector<int32_t> e; return e; }
        return it->second;
    }

    void finalize() {
        auto cnt = [&](const char* k) { auto& v = I(k); return v.empty() ? 0 : (int)v[0]; };
        n_basin = cnt("n_basin"); n_trc = cnt("n_trc"); n_pump = cnt("n_pump");
        n_outlet = cnt("n_outlet"); n_ud = cnt("n_ud"); n_lr = cnt("n_lr");
        n_manning = cnt("n_manning"); n_pid = cnt("n_pid"); n_cc = cnt("n_cc");
        n_lb = cnt("n_lb"); n_fb = cnt("n_fb"); n_prio = cnt("n_prio");
        n_dc = cnt("n_dc");
        dc_truth.assign(n_dc ? I("dc_cond_ptr")[n_dc] : 0, 0);
        dc_state.assign(n_dc, -1);
        n_ud_link = n_ud ? I("ud_link_ptr")[n_ud] : 0;
        // state_components order, core/src/parameter.jl:11-22
        off_trc = 0; off_pump = off_trc + n_trc; off_outlet = off_pump + n_pump;
        off_ud_in = off_outlet + n_outlet; off_ud_out = off_ud_in + n_ud_link;
        off_lr = off_ud_out + n_ud; off_manning = off_lr + n_lr; off_evap = off_manning + n_manning;
        off_infil = off_evap + n_basin; off_integral = off_infil + n_basin;
        n_state = off_integral + n_pid; n_reduced = n_basin + n_pid;
        if (!D("level_difference_threshold").empty())
            level_difference_threshold = D("level_difference_threshold")[0];
        // PCHIP tables
        auto& tp = I("tab_ptr");
        tabs.clear();
        int n_tab = tp.empty() ? 0 : (int)tp.size() - 1;
        for (int k = 0; k < n_tab; k++) {
            Pchip p;
            p.t.assign(D("tab_t").begin() + tp[k], D("tab_t").begin() + tp[k + 1]);
            p.u.assign(D("tab_u").begin() + tp[k], D("tab_u").begin() + tp[k + 1]);
            p.left_linear = I("tab_left_linear")[k] != 0;
            p.right_linear = I("tab_right_linear")[k] != 0;
            p.build();
            tabs.push_back(p);
        }
        auto ensure = [&](const char* k, size_t n, double v) {
            auto& x = Dm(k); if (x.size() != n) x.assign(n, v);
        };
        for (const char* k : {"precipitation", "surface_runoff", "potential_evaporation", "drainage",
                              "infiltration", "cumulative_precipitation",
                              "cumulative_surface_runoff", "cumulative_drainage"})
            ensure(k, n_basin, 0.0);
        ensure("fb_cumulative_flow", n_fb, 0.0);
        ensure("fb_incr", n_fb, 0.0);
        storage_prev = D("storage0");
        level_prev.assign(n_basin, 0.0);
        finalized = true;
        bind();
    }

    // get_state_index(state_ranges, id; inflow) — core/src/util.jl:878-895
    int state_of_node(int type, int idx, bool inflow) const {
        switch (type) {
            case TRC: return off_trc + idx;
            case PUMP: return off_pump + idx;
            case OUTLET: return off_outlet + idx;
            case LR: return off_lr + idx;
            case MANNING: return off_manning + idx;
            case USER_DEMAND: return inflow ? off_ud_in + idx : off_ud_out + idx;
            default: return -1;
        }
    }
    // state of the flow link (basin b -> node): link_to_state_idx lookup, util.jl:904-913
    int state_of_basin_outflow(int b, int type, int idx) const {
        if (type == USER_DEMAND) {
            auto lp = ud_link_ptr;
            for (int k = lp[idx]; k < lp[idx + 1]; k++)
                if (ud_link_src_type[k] == BASIN && ud_link_src_idx[k] == b)
                    return off_ud_in + k;
        }
        return state_of_node(type, idx, true);
    }

    // reduce_state! — core/src/util.jl:745-791
    template <class T> void reduce_state(const T* u, T* ur) const {
        auto ip = bin_ptr, it = bin_type, ii = bin_idx;
        auto op = bout_ptr, ot = bout_type, oi = bout_idx;
        for (int i = 0; i < n_basin; i++) {
            T acc(0.0);
            for (int k = ip[i]; k < ip[i + 1]; k++) {
                int st = state_of_node(it[k], ii[k], false);
                if (st < 0) continue;
                acc += u[st];
            }
            for (int k = op[i]; k < op[i + 1]; k++) {
                int st = state_of_basin_outflow(i, ot[k], oi[k]);
                if (st < 0) continue;
                acc -= u[st];
            }
            acc -= u[off_evap + i];
            acc -= u[off_infil + i];
            ur[i] = acc;
        }
        for (int k = 0; k < n_pid; k++) ur[n_basin + k] = u[off_integral + k];
    }
};

template <class T> struct Cache {
    std::vector<T> storage, factor, level, area, flow_rate_pump, flow_rate_outlet, error_pid;
    void resize(const Model& m) {
        storage.assign(m.n_basin, T(0.0)); factor = storage; level = storage; area = storage;
        flow_rate_pump.assign(m.n_pump, T(0.0)); flow_rate_outlet.assign(m.n_outlet, T(0.0));
        error_pid.assign(m.n_pid, T(0.0));
    }
};

// LinearInterpolationIntInv evaluation — third-party DataInterpolations (call site
// core/src/util.jl:42-44; construction core/src/read.jl:2281-2296).
template <class T> T storage_to_level(const Model& m, int b, T S) {
    auto& pp = m.prof_ptr;
    const double* h = m.prof_level + pp[b];
    const double* a = m.prof_dSdh + pp[b];
    const double* cs = m.prof_cumS + pp[b];
    int n = pp[b + 1] - pp[b];
    double Sv = val(S);
    if (Sv < cs[0]) return T(h[0]) + (S - T(cs[0])) / T(a[0]);
    if (Sv > cs[n - 1]) return T(h[n - 1]) + (S - T(cs[n - 1])) / T(a[n - 1]);
    int idx = int(std::upper_bound(cs, cs + n, Sv) - cs);
    idx = std::min(std::max(idx, 1), n - 1) - 1;
    T dS = S - T(cs[idx]);
    double x = a[idx];
    double slope = (a[idx + 1] - a[idx]) / (h[idx + 1] - h[idx]);
    // cancellation-free form of the quadratic root (exact for slope == 0); the reference builds
    // dS_dh by finite differences of the integrated profile (read.jl:2171-2185), which leaves
    // slopes like -5.7e-15 on constant-area segments
    return T(h[idx]) + T(2.0) * dS / (T(x) + xsqrt(T(x * x) + T(slope) * (T(2.0) * dS)));
}

// level_to_area: LinearInterpolation with Constant extrapolation (read.jl:2303-2315)
template <class T> T level_to_area(const Model& m, int b, T level) {
    auto& pp = m.prof_ptr;
    const double* h = m.prof_level + pp[b];
    const double* a = m.prof_area + pp[b];
    int n = pp[b + 1] - pp[b];
    double lv = val(level);
    if (lv < h[0]) return T(a[0]);
    if (lv > h[n - 1]) return T(a[n - 1]);
    int idx = int(std::upper_bound(h, h + n, lv) - h);
    idx = std::min(std::max(idx, 1), n - 1) - 1;
    double slope = (a[idx + 1] - a[idx]) / (h[idx + 1] - h[idx]);
    return T(a[idx]) + T(slope) * (level - T(h[idx]));
}

// get_level — core/src/util.jl:170-191
template <class T> T get_level(const Model& m, const Cache<T>& c, int type, int idx) {
    if (type == BASIN) return c.level[idx];
    if (type == LEVEL_BOUNDARY) return T(m.lb_level[idx]);
    return T(-std::numeric_limits<double>::infinity());  // Terminal
}
// get_low_storage_factor — core/src/util.jl:323-330
template <class T> T low_storage_factor(const Cache<T>& c, int type, int idx) {
    return type == BASIN ? c.factor[idx] : T(1.0);
}

// set_current_basin_properties! + formulate_storages! — core/src/solve.jl:120-204
template <class T> void basin_properties(const Model& m, Cache<T>& c, const T* ur, double t) {
    double dt = m.h_override >= 0.0 ? m.h_override : t - m.tprev;
    auto& s0 = m.storage0;
    auto& fa = m.fixed_area;
    std::vector<double> fb_cum(m.n_fb);
    for (int k = 0; k < m.n_fb; k++) fb_cum[k] = m.fb_cumulative_flow[k] + m.fb_incr[k];
    for (int i = 0; i < m.n_basin; i++) {
        double ccp = m.cumulative_precipitation[i] + fa[i] * m.precipitation[i] * dt;
        double ccr = m.cumulative_surface_runoff[i] + dt * m.surface_runoff[i];
        double ccd = m.cumulative_drainage[i] + dt * m.drainage[i];
        if (!m.merged_exact) {
            T S = T(s0[i]);
            S += ur[i];
            S += T(ccp); S += T(ccr); S += T(ccd);
            c.storage[i] = S;
        } else {
            c.storage[i] = T(0.0);  // filled below
        }
    }
    if (!m.merged_exact) {
        for (int k = 0; k < m.n_fb; k++)
            if (m.fb_dst_type[k] == BASIN) c.storage[m.fb_dst_idx[k]] += T(fb_cum[k]);
    } else {
        std::vector<double> exact(m.n_basin);
        for (int i = 0; i < m.n_basin; i++) {
            double dtl = dt;
            double ccp = m.cumulative_precipitation[i] + fa[i] * m.precipitation[i] * dtl;
            double ccr = m.cumulative_surface_runoff[i] + dtl * m.surface_runoff[i];
            double ccd = m.cumulative_drainage[i] + dtl * m.drainage[i];
            exact[i] = (ccp + ccr) + ccd;
        }
        for (int k = 0; k < m.n_fb; k++)
            if (m.fb_dst_type[k] == BASIN) exact[m.fb_dst_idx[k]] += fb_cum[k];
        for (int i = 0; i < m.n_basin; i++) c.storage[i] = (T(s0[i]) + ur[i]) + T(exact[i]);
    }
    for (int i = 0; i < m.n_basin; i++) {
        T S = c.storage[i];
        c.factor[i] = reduction_factor(S, m.low_storage_threshold[i]);
        c.level[i] = storage_to_level(m, i, S);
        c.area[i] = level_to_area(m, i, c.level[i]);
    }
}

// linear_resistance_flow — core/src/solve.jl:431-450
template <class T> T lr_flow(const Model& m, const Cache<T>& c, int k) {
    int st = m.lr_src_type[k], si = m.lr_src_idx[k];
    int dt = m.lr_dst_type[k], di = m.lr_dst_idx[k];
    T h_a = get_level(m, c, st, si), h_b = get_level(m, c, dt, di);
    T dh = h_a - h_b;
    T q_unlimited = dh / T(m.lr_resistance[k]);
    double qmax = m.lr_max_flow_rate[k];
    T q = jl_clamp(q_unlimited, T(-qmax), T(qmax));
    T f = val(q_unlimited) > 0 ? low_storage_factor(c, st, si) : low_storage_factor(c, dt, di);
    return q * f;
}

// manning_resistance_flow — core/src/solve.jl:528-589
template <class T> T manning_flow(const Model& m, const Cache<T>& c, int k) {
    int st = m.mn_src_type[k], si = m.mn_src_idx[k];
    int dt = m.mn_dst_type[k], di = m.mn_dst_idx[k];
    T h_a = get_level(m, c, st, si), h_b = get_level(m, c, dt, di);
    double bottom_a = m.mn_upstream_bottom[k], bottom_b = m.mn_downstream_bottom[k];
    double slope = m.mn_profile_slope[k], width = m.mn_profile_width[k];
    double n = m.mn_manning_n[k], L = m.mn_length[k];
    T d_a = h_a - T(bottom_a), d_b = h_b - T(bottom_b);
    T d = T(0.5) * (d_a + d_b);
    T A_a = T(width) * d + T(slope) * (d_a * d_a);
    T A_b = T(width) * d + T(slope) * (d_b * d_b);
    T A = T(0.5) * (A_a + A_b);
    double slope_unit_length = std::sqrt(slope * slope + 1.0);
    T P_a = T(width) + T(2.0) * d_a * T(slope_unit_length);
    T P_b = T(width) + T(2.0) * d_b * T(slope_unit_length);
    T R_h_a = A_a / P_a, R_h_b = A_b / P_b;
    T R_h = T(0.5) * (R_h_a + R_h_b);
    T dh = h_a - h_b;
    const double nu = 1.004e-6;
    T thr = T(2000 * nu) * T(n) * xcbrt(R_h) / A;
    thr = thr * thr;
    thr = jl_max(thr, 1.0e-5);
    T q = A / T(n) * xcbrt(R_h * R_h) * relaxed_root(dh / T(L), thr);
    T f = val(q) > 0 ? low_storage_factor(c, st, si) : low_storage_factor(c, dt, di);
    return q * f;
}

// tabulated_rating_curve_flow — core/src/solve.jl:452-474
template <class T> T trc_flow(const Model& m, const Cache<T>& c, int k) {
    int st = m.trc_src_type[k], si = m.trc_src_idx[k];
    int dt = m.trc_dst_type[k], di = m.trc_dst_idx[k];
    T h_a = get_level(m, c, st, si), h_b = get_level(m, c, dt, di);
    T dh = h_a - h_b;
    T factor = low_storage_factor(c, st, si);
    const Pchip& qh = m.tabs[m.trc_table_idx[k]];
    T q = factor * qh.eval(h_a);
    q *= reduction_factor(dh, m.level_difference_threshold);
    double max_down = m.trc_max_downstream_level[k];
    q *= reduction_factor(T(max_down) - h_b, m.level_difference_threshold);
    return q;
}

// formulate_pump_or_outlet_flow! — core/src/solve.jl:657-770
template <class T>
void pump_outlet_flows(const Model& m, const Cache<T>& c, T* du, bool outlet, int relevant) {
    int n = outlet ? m.n_outlet : m.n_pump;
    int off = outlet ? m.off_outlet : m.off_pump;
    const int32_t* ct = outlet ? m.outlet_control_type : m.pump_control_type;
    const int32_t* srct = outlet ? m.outlet_src_type : m.pump_src_type;
    const int32_t* srci = outlet ? m.outlet_src_idx : m.pump_src_idx;
    const int32_t* dstt = outlet ? m.outlet_dst_type : m.pump_dst_type;
    const int32_t* dsti = outlet ? m.outlet_dst_idx : m.pump_dst_idx;
    const double* fr = outlet ? m.outlet_flow_rate : m.pump_flow_rate;
    const double* minq = outlet ? m.outlet_min_flow_rate : m.pump_min_flow_rate;
    const double* maxq = outlet ? m.outlet_max_flow_rate : m.pump_max_flow_rate;
    const double* minup = outlet ? m.outlet_min_upstream_level : m.pump_min_upstream_level;
    const double* maxdn = outlet ? m.outlet_max_downstream_level : m.pump_max_downstream_level;
    const double* fdem = outlet ? m.outlet_flow_demand : m.pump_flow_demand;  // NaN: none attached
    const std::vector<T>& cur = outlet ? c.flow_rate_outlet : c.flow_rate_pump;
    for (int k = 0; k < n; k++) {
        if (ct[k] != relevant) continue;
        T flow_rate = ct[k] != CONTROL_NONE ? cur[k] : T(fr[k]);
        T src_level = get_level(m, c, srct[k], srci[k]);
        T dst_level = get_level(m, c, dstt[k], dsti[k]);
        T q = flow_rate * low_storage_factor(c, srct[k], srci[k]);
        double lower = minq[k], upper = maxq[k];
        if (fdem && fdem[k] == fdem[k]) {  // allocation inactive: demand is a lower bound
            double td = fdem[k];
            lower = td > upper ? upper : (td < lower ? lower : td);
        }
        q = jl_clamp(q, T(lower), T(upper));
        if (outlet) q *= reduction_factor(src_level - dst_level, m.level_difference_threshold);
        q *= reduction_factor(src_level - T(minup[k]), m.level_difference_threshold);
        q *= reduction_factor(T(maxdn[k]) - dst_level, m.level_difference_threshold);
        du[off + k] = q;
    }
}

// formulate_flow!(UserDemand) — core/src/solve.jl:348-404
template <class T> void user_demand_flows(const Model& m, const Cache<T>& c, T* du) {
    auto& lp = m.ud_link_ptr;
    for (int i = 0; i < m.n_ud; i++) {
        double q_total_demand = 0.0;
        for (int p = 0; p < m.n_prio; p++) {
            if (!m.ud_has_priority[i * m.n_prio + p]) continue;
            q_total_demand += std::min(m.ud_allocated[i * m.n_prio + p],
                                       m.ud_demand[i * m.n_prio + p]);
        }
        int n_links = lp[i + 1] - lp[i];
        double equal_split = n_links == 0 ? 0.0 : q_total_demand / n_links;
        T q_total_actual(0.0);
        for (int k = lp[i]; k < lp[i + 1]; k++) {
            int st = m.ud_link_src_type[k], si = m.ud_link_src_idx[k];
            T f_low = low_storage_factor(c, st, si);
            T source_level = get_level(m, c, st, si);
            T f_red = reduction_factor(source_level - T(m.ud_min_level[i]),
                                       m.level_difference_threshold);
            double alloc = m.ud_link_alloc[k];
            double q_target = std::isinf(alloc) ? equal_split : alloc;
            T q_k = T(q_target) * f_low * f_red;
            du[m.off_ud_in + k] = q_k;
            q_total_actual += q_k;
        }
        du[m.off_ud_out + i] = q_total_actual * T(m.ud_return_factor[i]);
    }
}

// formulate_flows! — core/src/solve.jl:811-834
template <class T> void formulate_flows(const Model& m, const Cache<T>& c, T* du, int control) {
    pump_outlet_flows(m, c, du, false, control);
    pump_outlet_flows(m, c, du, true, control);
    if (control == CONTROL_NONE) {
        for (int k = 0; k < m.n_lr; k++) du[m.off_lr + k] = lr_flow(m, c, k);
        for (int k = 0; k < m.n_manning; k++) du[m.off_manning + k] = manning_flow(m, c, k);
        for (int k = 0; k < m.n_trc; k++) du[m.off_trc + k] = trc_flow(m, c, k);
        user_demand_flows(m, c, du);
    }
}

// formulate_continuous_control! — core/src/solve.jl:101-113; compound_variable_value —
// core/src/callback.jl:757-763
template <class T> void continuous_control(const Model& m, Cache<T>& c, const T* du) {
    auto& sp = m.cc_sv_ptr;
    for (int i = 0; i < m.n_cc; i++) {
        T value(0.0);
        for (int j = sp[i]; j < sp[i + 1]; j++) {
            int kind = m.cc_sv_kind[j], idx = m.cc_sv_idx[j];
            T x;
            if (kind == SV_BASIN_LEVEL) x = c.level[idx];
            else if (kind == SV_BASIN_STORAGE) x = c.storage[idx];
            else if (kind == SV_DU_STATE) x = du[idx];
            else x = T(m.cc_sv_const[j]);
            value += T(m.cc_sv_weight[j]) * x;
        }
        T out = m.tabs[m.cc_table_idx[i]].eval(value);
        if (m.cc_target_is_outlet[i]) c.flow_rate_outlet[m.cc_target_idx[i]] = out;
        else c.flow_rate_pump[m.cc_target_idx[i]] = out;
    }
}

// formulate_dstorage_wrt_time — core/src/solve.jl:321-346 (get_flow: core/src/graph.jl:346-366)
template <class T> T dstorage_wrt_time(const Model& m, const T* du, int b) {
    auto &ip = m.bin_ptr, &it = m.bin_type, &ii = m.bin_idx;
    auto &op = m.bout_ptr, &ot = m.bout_type, &oi = m.bout_idx;
    T ds(0.0);
    for (int k = ip[b]; k < ip[b + 1]; k++) {
        if (it[k] == FLOW_BOUNDARY) { ds += T(m.fb_rate_now[ii[k]]); continue; }
        int st = m.state_of_node(it[k], ii[k], false);
        ds += du[st];
    }
    for (int k = op[b]; k < op[b + 1]; k++) {
        int st = m.state_of_basin_outflow(b, ot[k], oi[k]);
        ds -= du[st];
    }
    ds += T(m.fixed_area[b] * m.precipitation[b]);
    ds += T(m.surface_runoff[b]);
    ds += T(m.drainage[b]);
    ds -= du[m.off_evap + b];
    ds -= du[m.off_infil + b];
    return ds;
}

// formulate_pid_control! — core/src/solve.jl:229-316
template <class T> void pid_control(const Model& m, Cache<T>& c, T* du, const T* ur) {
    for (int i = 0; i < m.n_pid; i++) {
        int L = m.pid_listen_basin[i];
        c.error_pid[i] = T(m.pid_target[i]) - c.level[L];
    }
    for (int i = 0; i < m.n_pid; i++) {
        du[m.off_integral + i] = c.error_pid[i];
        int L = m.pid_listen_basin[i];
        T flow_rate(0.0);
        double K_p = m.pid_proportional[i], K_i = m.pid_integral[i],
               K_d = m.pid_derivative[i];
        T D(1.0), area(1.0);
        if (K_d != 0.0) { area = c.area[L]; D = T(1.0) - T(K_d) / area; }
        if (K_p != 0.0) flow_rate += T(K_p) * c.error_pid[i] / D;
        if (K_i != 0.0) flow_rate += T(K_i) * ur[m.n_basin + i] / D;
        if (K_d != 0.0) {
            double dtarget = 0.0;  // target is a ConstantInterpolation (solve.jl:298-300)
            T ds_old = dstorage_wrt_time(m, du, L);
            flow_rate += T(K_d) * (T(dtarget) - ds_old / area) / D;
        }
        if (m.pid_target_is_outlet[i]) c.flow_rate_outlet[m.pid_target_idx[i]] = flow_rate;
        else c.flow_rate_pump[m.pid_target_idx[i]] = flow_rate;
    }
}

// water_balance!(du, u_reduced, ...) — core/src/solve.jl:35-84
template <class T> void water_balance(const Model& m, Cache<T>& c, const T* ur, double t, T* du) {
    for (int k = 0; k < m.n_state; k++) du[k] = T(0.0);
    basin_properties(m, c, ur, t);
    // update_vertical_flux! — core/src/solve.jl:209-227
    for (int i = 0; i < m.n_basin; i++) {
        du[m.off_evap + i] = c.area[i] * c.factor[i] * T(m.potential_evaporation[i]);
        du[m.off_infil + i] = c.factor[i] * T(m.infiltration[i]);
    }
    formulate_flows(m, c, du, CONTROL_NONE);
    continuous_control(m, c, du);
    formulate_flows(m, c, du, CONTROL_CONTINUOUS);
    pid_control(m, c, du, ur);
    formulate_flows(m, c, du, CONTROL_PID);
}

// ------------------------------------------------------------------------------------------
// Jacobian by forward-mode dual numbers, one column (or one colour) per pass —
// core/src/differentiation.jl:361-390
// ------------------------------------------------------------------------------------------
void jac_dense(const Model& m, const double* ur, double t, double* J /* n_state x n_r, col-major */) {
    Cache<Dual> c; c.resize(m);
    std::vector<Dual> x(m.n_reduced), du(m.n_state);
    for (int j = 0; j < m.n_reduced; j++) {
        for (int k = 0; k < m.n_reduced; k++) x[k] = Dual(ur[k], k == j ? 1.0 : 0.0);
        water_balance(m, c, x.data(), t, du.data());
        for (int r = 0; r < m.n_state; r++) J[(size_t)j * m.n_state + r] = du[r].d;
    }
}

void build_coloring(Model& m) {
    // greedy colouring of the column intersection graph of the given CSC pattern
    int nr = m.n_reduced;
    std::vector<std::vector<int>> rows_of_col(nr), cols_of_row(m.n_state);
    for (size_t k = 0; k < m.pat_rows.size(); k++) {
        rows_of_col[m.pat_cols[k]].push_back(m.pat_rows[k]);
        cols_of_row[m.pat_rows[k]].push_back(m.pat_cols[k]);
    }
    m.color_of_col.assign(nr, -1);
    m.n_colors = 0;
    std::vector<int> mark(nr + 1, -1);
    for (int j = 0; j < nr; j++) {
        for (int r : rows_of_col[j])
            for (int j2 : cols_of_row[r])
                if (m.color_of_col[j2] >= 0) mark[m.color_of_col[j2]] = j;
        int col = 0;
        while (mark[col] == j) col++;
        m.color_of_col[j] = col;
        m.n_colors = std::max(m.n_colors, col + 1);
    }
}

// values of the pattern entries (CSC order) by coloured dual passes
void jac_colored(const Model& m, const double* ur, double t, double* vals) {
    Cache<Dual> c; c.resize(m);
    std::vector<Dual> x(m.n_reduced), du(m.n_state);
    for (int col = 0; col < m.n_colors; col++) {
        for (int k = 0; k < m.n_reduced; k++) x[k] = Dual(ur[k], m.color_of_col[k] == col ? 1.0 : 0.0);
        water_balance(m, c, x.data(), t, du.data());
        for (size_t k = 0; k < m.pat_rows.size(); k++)
            if (m.color_of_col[m.pat_cols[k]] == col) vals[k] = du[m.pat_rows[k]].d;
    }
}

// calc_J_inner! (differentiation.jl:153-262) + jacobian2W!: W = A*J_i - I/gamma from the CSC
// pattern values.  `at(i, j)` returns a reference to W[i, j] (dense or sparse storage).
template <class At> void assemble_W(const Model& m, const double* vals, double gamma, At&& at) {
    int nr = m.n_reduced;
    for (size_t k = 0; k < m.pat_rows.size(); k++) {
        int row = m.pat_rows[k], col = m.pat_cols[k];
        double v = vals[k];
        auto conn = [&](const int32_t* st, const int32_t* si, const int32_t* dt, const int32_t* di,
                        int idx) {
            if (st[idx] == BASIN) at(si[idx], col) -= v;
            if (dt[idx] == BASIN) at(di[idx], col) += v;
        };
        if (row < m.off_pump) conn(m.trc_src_type, m.trc_src_idx, m.trc_dst_type, m.trc_dst_idx, row - m.off_trc);
        else if (row < m.off_outlet) conn(m.pump_src_type, m.pump_src_idx, m.pump_dst_type, m.pump_dst_idx, row - m.off_pump);
        else if (row < m.off_ud_in) conn(m.outlet_src_type, m.outlet_src_idx, m.outlet_dst_type, m.outlet_dst_idx, row - m.off_outlet);
        else if (row < m.off_ud_out) {
            int k2 = row - m.off_ud_in;
            if (m.ud_link_src_type[k2] == BASIN) at(m.ud_link_src_idx[k2], col) -= v;
        } else if (row < m.off_lr) {
            int i = row - m.off_ud_out;
            if (m.ud_dst_type[i] == BASIN) at(m.ud_dst_idx[i], col) += v;
        } else if (row < m.off_manning) conn(m.lr_src_type, m.lr_src_idx, m.lr_dst_type, m.lr_dst_idx, row - m.off_lr);
        else if (row < m.off_evap) conn(m.mn_src_type, m.mn_src_idx, m.mn_dst_type, m.mn_dst_idx, row - m.off_manning);
        else if (row < m.off_infil) at(row - m.off_evap, col) -= v;
        else if (row < m.off_integral) at(row - m.off_infil, col) -= v;
        else at(m.n_basin + (row - m.off_integral), col) += v;
    }
    double inv = 1.0 / gamma;
    for (int i = 0; i < nr; i++) at(i, i) = at(i, i) - inv;
}

void calc_W_inner(const Model& m, const double* vals, double gamma, double* W) {
    int nr = m.n_reduced;
    std::fill(W, W + (size_t)nr * nr, 0.0);
    assemble_W(m, vals, gamma, [&](int i, int j) -> double& { return W[(size_t)j * nr + i]; });
}

// dense LU with partial pivoting (stands in for KLU; any correct LU agrees to rounding)
bool dense_factor(int n, std::vector<double>& A /* col-major, overwritten by L\\U */,
                  std::vector<int>& piv) {
    piv.resize(n);
    for (int k = 0; k < n; k++) {
        int p = k; double best = std::fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++) {
            double v = std::fabs(A[(size_t)k * n + i]);
            if (v > best) { best = v; p = i; }
        }
        if (best == 0.0 || best != best) return false;
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < n; j++) std::swap(A[(size_t)j * n + k], A[(size_t)j * n + p]);
        double d = A[(size_t)k * n + k];
        for (int i = k + 1; i < n; i++) {
            double l = A[(size_t)k * n + i] / d;
            A[(size_t)k * n + i] = l;
            if (l == 0.0) continue;
            for (int j = k + 1; j < n; j++) A[(size_t)j * n + i] -= l * A[(size_t)j * n + k];
        }
    }
    return true;
}
void dense_apply(int n, const std::vector<double>& A, const std::vector<int>& piv, double* b) {
    for (int k = 0; k < n; k++)
        if (piv[k] != k) std::swap(b[k], b[piv[k]]);
    for (int k = 0; k < n; k++)
        for (int i = k + 1; i < n; i++) b[i] -= A[(size_t)k * n + i] * b[k];
    for (int i = n - 1; i >= 0; i--) {
        double v = b[i];
        for (int j = i + 1; j < n; j++) v -= A[(size_t)j * n + i] * b[j];
        b[i] = v / A[(size_t)i * n + i];
    }
}
bool dense_solve(int n, std::vector<double>& A /* col-major, destroyed */, double* b) {
    std::vector<int> piv;
    if (!dense_factor(n, A, piv)) return false;
    dense_apply(n, A, piv, b);
    return true;
}

// min_low_storage_factor / min_low_user_demand_level_factor — core/src/util.jl:941-981
double min_low_storage_factor(const Model& m, const std::vector<double>& storage_now, int type, int idx) {
    if (type != BASIN) return 1.0;
    double thr = m.low_storage_threshold[idx];
    return reduction_factor(std::min(storage_now[idx], m.storage_prev[idx]) - 2 * thr, thr);
}

// limit_flow! — core/src/solve.jl:840-984
void limit_flow(const Model& m, double* u, const double* uprev, double dt, double t) {
    Cache<double> c; c.resize(m);
    std::vector<double> ur(m.n_reduced);
    m.reduce_state(u, ur.data());
    basin_properties(m, c, ur.data(), t);
    auto lim = [&](int st, double lo, double hi) {
        double a = uprev[st] + lo * dt, b = uprev[st] + hi * dt;
        u[st] = jl_clamp(u[st], a, b);
    };
    const double inf = std::numeric_limits<double>::infinity();
    for (int k = 0; k < m.n_trc; k++) lim(m.off_trc + k, 0.0, inf);
    for (int k = 0; k < m.n_pump; k++)
        lim(m.off_pump + k, m.pump_min_flow_rate[k], m.pump_max_flow_rate[k]);
    for (int k = 0; k < m.n_outlet; k++)
        lim(m.off_outlet + k, m.outlet_min_flow_rate[k], m.outlet_max_flow_rate[k]);
    for (int k = 0; k < m.n_lr; k++)
        lim(m.off_lr + k, -m.lr_max_flow_rate[k], m.lr_max_flow_rate[k]);
    auto& lp = m.ud_link_ptr;
    for (int i = 0; i < m.n_ud; i++) {
        bool from_ts = m.ud_demand_from_timeseries[i] != 0;
        double allocated_total = 0.0;
        if (!from_ts)
            for (int p = 0; p < m.n_prio; p++)
                allocated_total += std::min(m.ud_demand_static[i * m.n_prio + p],
                                            m.ud_allocated[i * m.n_prio + p]);
        int n_links = lp[i + 1] - lp[i];
        double equal_split = n_links == 0 ? 0.0 : allocated_total / n_links;
        for (int k = lp[i]; k < lp[i + 1]; k++) {
            double alloc = m.ud_link_alloc[k];
            double q_k_max = std::isinf(alloc) ? equal_split : alloc;
            double lo, hi;
            if (from_ts) { lo = 0.0; hi = inf; }
            else {
                int st = m.ud_link_src_type[k], si = m.ud_link_src_idx[k];
                double fb = min_low_storage_factor(m, c.storage, st, si);
                double fl = 1.0;
                if (st == BASIN)
                    fl = reduction_factor(std::min(c.level[si], m.level_prev[si]) -
                                              m.ud_min_level[i] - 2 * m.level_difference_threshold,
                                          m.level_difference_threshold);
                lo = fb * fl * q_k_max; hi = q_k_max;
            }
            lim(m.off_ud_in + k, lo, hi);
        }
    }
    for (int i = 0; i < m.n_basin; i++) {
        double fmin = min_low_storage_factor(m, c.storage, BASIN, i);
        double infil = m.infiltration[i];
        lim(m.off_evap + i, 0.0, inf);
        lim(m.off_infil + i, fmin * infil, infil);
    }
}

// ------------------------------------------------------------------------------------------
// DiscreteControl — apply_discrete_control! / set_new_control_state! / set_control_params!
// (core/src/callback.jl:608-786).  `flow` holds du-like values per state (the RHS at t = 0, the
// mean flow (u - uprev)/h after an accepted implicit-Euler step: what get_du returns).
// ------------------------------------------------------------------------------------------
void dc_write_params(Model& m) {
    for (int k = 0; k < m.n_pump; k++) {
        int d = m.pump_dc ? m.pump_dc[k] : -1;
        if (d < 0 || m.dc_state[d] < 0) continue;
        int row = m.pump_cs_ptr[k] + m.dc_state[d];
        m.pump_flow_rate[k] = m.pump_cs_flow_rate[row];
        m.pump_min_flow_rate[k] = m.pump_cs_min_flow_rate[row];
        m.pump_max_flow_rate[k] = m.pump_cs_max_flow_rate[row];
        m.pump_min_upstream_level[k] = m.pump_cs_min_upstream_level[row];
        m.pump_max_downstream_level[k] = m.pump_cs_max_downstream_level[row];
    }
    for (int k = 0; k < m.n_outlet; k++) {
        int d = m.outlet_dc ? m.outlet_dc[k] : -1;
        if (d < 0 || m.dc_state[d] < 0) continue;
        int row = m.outlet_cs_ptr[k] + m.dc_state[d];
        m.outlet_flow_rate[k] = m.outlet_cs_flow_rate[row];
        m.outlet_min_flow_rate[k] = m.outlet_cs_min_flow_rate[row];
        m.outlet_max_flow_rate[k] = m.outlet_cs_max_flow_rate[row];
        m.outlet_min_upstream_level[k] = m.outlet_cs_min_upstream_level[row];
        m.outlet_max_downstream_level[k] = m.outlet_cs_max_downstream_level[row];
    }
    for (int k = 0; k < m.n_lr; k++) {
        int d = m.lr_dc ? m.lr_dc[k] : -1;
        if (d < 0 || m.dc_state[d] < 0) continue;
        int row = m.lr_cs_ptr[k] + m.dc_state[d];
        m.lr_resistance[k] = m.lr_cs_resistance[row];
        m.lr_max_flow_rate[k] = m.lr_cs_max_flow_rate[row];
    }
    for (int k = 0; k < m.n_manning; k++) {
        int d = m.mn_dc ? m.mn_dc[k] : -1;
        if (d < 0 || m.dc_state[d] < 0) continue;
        int row = m.mn_cs_ptr[k] + m.dc_state[d];
        m.mn_length[k] = m.mn_cs_length[row];
        m.mn_manning_n[k] = m.mn_cs_manning_n[row];
        m.mn_profile_width[k] = m.mn_cs_profile_width[row];
        m.mn_profile_slope[k] = m.mn_cs_profile_slope[row];
    }
    for (int k = 0; k < m.n_pid; k++) {
        int d = m.pid_dc ? m.pid_dc[k] : -1;
        if (d < 0 || m.dc_state[d] < 0) continue;
        int row = m.pid_cs_ptr[k] + m.dc_state[d];
        m.pid_target[k] = m.pid_cs_target[row];
        m.pid_proportional[k] = m.pid_cs_proportional[row];
        m.pid_integral[k] = m.pid_cs_integral[row];
        m.pid_derivative[k] = m.pid_cs_derivative[row];
    }
    for (int k = 0; k < m.n_trc; k++) {
        int d = m.trc_dc ? m.trc_dc[k] : -1;
        if (d < 0 || m.dc_state[d] < 0) continue;
        const_cast<int32_t*>(m.trc_table_idx)[k] = m.trc_cs_table[m.trc_cs_ptr[k] + m.dc_state[d]];
    }
}

void apply_discrete_control(Model& m, const Cache<double>& c, const double* flow, bool initial) {
    bool any = false;
    for (int i = 0; i < m.n_dc; i++) {
        int c0 = m.dc_cond_ptr[i], c1 = m.dc_cond_ptr[i + 1];
        int bits = 0;
        bool change = false;
        for (int cd = c0; cd < c1; cd++) {
            double value = 0.0;
            for (int j = m.dc_csv_ptr[cd]; j < m.dc_csv_ptr[cd + 1]; j++) {
                int kind = m.dc_sv_kind[j], idx = m.dc_sv_idx[j];
                double x;
                if (kind == SV_BASIN_LEVEL) x = c.level[idx];
                else if (kind == SV_BASIN_STORAGE) x = c.storage[idx];
                else if (kind == SV_DU_STATE) x = flow[idx];
                else x = m.dc_sv_const[j];
                value += m.dc_sv_weight[j] * x;
            }
            int told = m.dc_truth[cd];
            int tnew = told ? (value >= m.dc_thr_lo[cd]) : (value > m.dc_thr_hi[cd]);
            if (tnew != told) { change = true; m.dc_truth[cd] = tnew; }
            bits |= tnew << (cd - c0);
        }
        if (initial || change) {
            int snew = m.dc_logic[m.dc_logic_ptr[i] + bits];
            if (snew < 0) m.dc_undefined++;
            else if (snew != m.dc_state[i]) {
                m.dc_state[i] = snew; m.dc_changes++; any = true;
                m.dc_record.push_back(i); m.dc_record.push_back(bits); m.dc_record.push_back(snew);
            }
        }
    }
    if (any) dc_write_params(m);
}

// ------------------------------------------------------------------------------------------
// Implicit Euler with NLNewton (OrdinaryDiffEq, third-party; semantics as recalled in SURVEY
// §8(c)): z0 = 0, gamma = 1, W = J - I/h, solve W dz = (h f(uprev+z) - z)/h, z <- z - dz,
// kappa = 1/100, max_iter = 10.  `full_newton = 0` evaluates J and W once per (sub-)step at
// (uprev, t + h) like the reference's nlsolve!; `full_newton = 1` refreshes them every
// iteration.
//
// One outer step [t0, t0 + dt] is covered by sub-steps: a sub-step whose Newton iteration
// fails, or that ends with negative storage (isoutofdomain, core/src/util.jl:918-924), is
// rejected and retried with half the length (what the reference's adaptive controller does on
// a failed nonlinear solve); after an accepted sub-step that converged within `grow_iters`
// iterations the nominal length doubles again, capped at dt.  With max_halvings = 0 this is
// plain fixed-step implicit Euler.
// ------------------------------------------------------------------------------------------
struct AdvanceOpts {
    double abstol = 1e-5, reltol = 1e-5;
    int max_iter = 10, full_newton = 0, grow_iters = 4, max_halvings = 0, predictor = 0;
    int early_exit = 0;   // 1: give up an attempt predicted not to converge within max_iter
    int shrink_log2 = 1;  // a rejected attempt is retried with h / 2^shrink_log2
};
struct StepStats { int iters = 0, converged = 0, negative_storage = 0, substeps = 0, rejects = 0; double ndz = 0; };

struct NewtonWork {
    std::vector<double> z, ustep, k, b, dz, ur, br, jv, W;
    std::vector<int> piv;
    Cache<double> c;
};

// factorise W = A J - I/h with J at `ur`; returns false if singular
static bool refresh_W(Model& m, NewtonWork& w, const double* ur, double tn, double h, bool sparse) {
    int nr = m.n_reduced;
    jac_colored(m, ur, tn, w.jv.data());
    if (sparse) {
        if (!m.lu_ready) {
            std::vector<std::vector<int>> pattern(nr);
            double dummy = 0.0;
            std::vector<double> ones(m.pat_rows.size(), 1.0);
            assemble_W(m, ones.data(), 1.0, [&](int i, int j) -> double& {
                pattern[i].push_back(j); dummy = 0.0; return dummy; });
            m.lu.analyse(nr, pattern);
            m.lu_ready = true;
        }
        m.lu.clear();
        assemble_W(m, w.jv.data(), h, [&](int i, int j) -> double& {
            return m.lu.at(m.lu.pos_of[i], m.lu.pos_of[j]); });
        return m.lu.factor();
    }
    w.W.assign((size_t)nr * nr, 0.0);
    calc_W_inner(m, w.jv.data(), h, w.W.data());
    return dense_factor(nr, w.W, w.piv);
}

// one nlsolve! attempt from uprev over a sub-step of length h ending at tn; z is the result
static int newton_attempt(Model& m, NewtonWork& w, const double* uprev, double tn, double h,
                          const AdvanceOpts& o, double* eta_io, int* iters_out, double* ndz_out) {
    int n = m.n_state, nr = m.n_reduced;
    bool sparse = nr > 48;
    // initial guess: z = 0 (OrdinaryDiffEq `extrapolant = :constant`) or the last accepted
    // increment rescaled to this sub-step (constant-flow predictor)
    if (o.predictor && m.hlast > 0.0 && (int)m.zlast.size() == n) {
        double r = h / m.hlast;
        for (int i = 0; i < n; i++) w.z[i] = m.zlast[i] * r;
    } else {
        std::fill(w.z.begin(), w.z.end(), 0.0);
    }
    const double kappa = 0.01;
    double eta = std::pow(std::max(*eta_io, 2.220446049250313e-16), 0.8);
    double ndz = 0, ndz_prev = 0;
    int status = 0, iter = 0;
    for (iter = 1; iter <= o.max_iter; iter++) {
        for (int i = 0; i < n; i++) w.ustep[i] = uprev[i] + w.z[i];
        m.reduce_state(w.ustep.data(), w.ur.data());
        if (iter == 1 || o.full_newton)
            if (!refresh_W(m, w, w.ur.data(), tn, h, sparse)) { status = -1; break; }
        water_balance(m, w.c, w.ur.data(), tn, w.k.data());
        for (int i = 0; i < n; i++) w.b[i] = (h * w.k[i] - w.z[i]) * (1.0 / h);
        m.reduce_state(w.b.data(), w.br.data());
        if (sparse) m.lu.solve(w.br.data());
        else dense_apply(nr, w.W, w.piv, w.br.data());
        // linu = gamma * (J_i c - b)
        std::fill(w.dz.begin(), w.dz.end(), 0.0);
        for (size_t e = 0; e < m.pat_rows.size(); e++)
            w.dz[m.pat_rows[e]] += w.jv[e] * w.br[m.pat_cols[e]];
        for (int i = 0; i < n; i++) w.dz[i] = (w.dz[i] - w.b[i]) * h;
        double acc = 0;
        for (int i = 0; i < n; i++) {
            double sc = o.abstol + std::max(std::fabs(uprev[i]), std::fabs(w.ustep[i])) * o.reltol;
            double r = w.dz[i] / sc;
            acc += r * r;
        }
        ndz_prev = ndz;
        ndz = std::sqrt(acc / n);
        for (int i = 0; i < n; i++) w.z[i] -= w.dz[i];
        if (!std::isfinite(ndz)) { status = -1; break; }
        double theta = 0;
        if (iter > 1) {
            theta = ndz / ndz_prev;
            if (theta > 2) { status = -1; break; }
            // Hairer-Wanner: the remaining iterations cannot bring eta*ndz below kappa
            if (o.early_exit && (theta >= 1.0 ||
                ndz * std::pow(theta, o.max_iter - iter) > kappa * (1.0 - theta))) { status = -1; break; }
        }
        if (iter > 1) eta = theta / (1 - theta);
        if ((iter == 1 && ndz < 1e-5) || (iter > 1 && eta >= 0 && eta * ndz < kappa)) { status = 1; break; }
    }
    if (iter > o.max_iter) iter = o.max_iter;
    *eta_io = eta;
    *iters_out = iter;
    *ndz_out = ndz;
    return status == 1;
}

int advance(Model& m, double* u, double t0, double dt, const AdvanceOpts& o, double* h_try,
            double* eta_old, StepStats* stats) {
    int n = m.n_state, nr = m.n_reduced;
    NewtonWork w;
    w.z.assign(n, 0.0); w.ustep.resize(n); w.k.resize(n); w.b.resize(n); w.dz.resize(n);
    w.ur.resize(nr); w.br.resize(nr); w.jv.resize(m.pat_rows.size()); w.c.resize(m);
    std::vector<double> uprev(n), uprop(n);
    std::vector<double> fb_full(m.fb_incr, m.fb_incr + m.n_fb);
    bool have_rate = m.fb_rate_step != nullptr && m.n_fb > 0;
    if (m.n_dc > 0) dc_write_params(m);  // the host may have refreshed the shared parameters
    double h_min = std::ldexp(dt, -o.max_halvings);
    double h_nom = std::min(*h_try > 0 ? *h_try : dt, dt);
    double elapsed = 0.0;
    StepStats st;
    int failed_members_steps = 0;
    while (elapsed < dt) {
        double rem = dt - elapsed;
        bool last = h_nom >= rem;
        double h = last ? rem : h_nom;
        double tn = last ? t0 + dt : t0 + (elapsed + h);
        for (int kf = 0; kf < m.n_fb; kf++)
            m.fb_incr[kf] = have_rate ? m.fb_rate_step[kf] * h : fb_full[kf] * (h / dt);
        m.h_override = h;
        std::copy(u, u + n, uprev.begin());
        int iters = 0; double ndz = 0;
        int conv = newton_attempt(m, w, uprev.data(), tn, h, o, eta_old, &iters, &ndz);
        st.iters += iters; st.ndz = ndz;
        for (int i = 0; i < n; i++) uprop[i] = uprev[i] + w.z[i];
        limit_flow(m, uprop.data(), uprev.data(), h, tn);
        // isoutofdomain / check_negative_storage at the proposed state
        m.reduce_state(uprop.data(), w.ur.data());
        basin_properties(m, w.c, w.ur.data(), tn);
        int neg = 0;
        for (int i = 0; i < m.n_basin; i++) if (w.c.storage[i] < 0) neg++;
        bool can_halve = o.max_halvings > 0 && 0.5 * h >= h_min * (1.0 - 1e-12);
        if ((conv && neg == 0) || !can_halve) {
            std::copy(uprop.begin(), uprop.end(), u);
            m.zlast.resize(n);
            for (int i = 0; i < n; i++) m.zlast[i] = uprop[i] - uprev[i];
            m.hlast = h;
            if (m.n_dc > 0) {
                std::vector<double> flow(n);
                for (int i = 0; i < n; i++) flow[i] = m.zlast[i] / h;
                apply_discrete_control(m, w.c, flow.data(), false);
            }
            // update_cumulative_flows! — core/src/callback.jl:125-182
            auto& fa = m.fixed_area;
            for (int i = 0; i < m.n_basin; i++) {
                m.cumulative_precipitation[i] += fa[i] * m.precipitation[i] * h;
                m.cumulative_surface_runoff[i] += h * m.surface_runoff[i];
                m.cumulative_drainage[i] += h * m.drainage[i];
            }
            for (int kf = 0; kf < m.n_fb; kf++) m.fb_cumulative_flow[kf] += m.fb_incr[kf];
            m.tprev = tn;
            elapsed = last ? dt : elapsed + h;
            st.substeps++;
            st.negative_storage += neg;
            if (!conv) failed_members_steps++;
            if (conv && iters <= o.grow_iters) h_nom = std::min(2.0 * h_nom, dt);
        } else {
            h_nom = std::ldexp(h, -o.shrink_log2);
            if (h_nom < h_min) h_nom = h_min;
            *eta_old = 1.0;
            st.rejects++;
        }
    }
    m.h_override = -1.0;
    for (int kf = 0; kf < m.n_fb; kf++) m.fb_incr[kf] = fb_full[kf];
    *h_try = h_nom;
    st.converged = failed_members_steps == 0;
    if (stats) *stats = st;
    return failed_members_steps == 0 ? 0 : 1;
}

// plain fixed-step implicit Euler (no sub-stepping): one nlsolve!, limiter, bookkeeping
int newton_step(Model& m, double* u, double t, double dt, double abstol, double reltol,
                int max_iter, double* eta_old, StepStats* stats, bool) {
    AdvanceOpts o; o.abstol = abstol; o.reltol = reltol; o.max_iter = max_iter;
    double h_try = dt;
    return advance(m, u, t, dt, o, &h_try, eta_old, stats);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C interface (ctypes)
// ------------------------------------------------------------------------------------------
extern "C" {

void* orc_create() { return new Model(); }
void orc_destroy(void* h) { delete (Model*)h; }
void* orc_clone(void* h) { return new Model(*(Model*)h); }
void orc_set_f64(void* h, const char* key, const double* p, int n) {
    ((Model*)h)->s.d[key].assign(p, p + n);
    ((Model*)h)->bind();
}
void orc_set_i32(void* h, const char* key, const int32_t* p, int n) {
    ((Model*)h)->s.i[key].assign(p, p + n);
    ((Model*)h)->bind();
}
int orc_get_f64(void* h, const char* key, double* p, int n) {
    auto& v = ((Model*)h)->s.d[key];
    if ((int)v.size() != n) return 1;
    std::copy(v.begin(), v.end(), p);
    return 0;
}
void orc_finalize(void* h) { ((Model*)h)->finalize(); }
void orc_set_pattern(void* h, const int32_t* rows, const int32_t* cols, int nnz) {
    Model& m = *(Model*)h;
    m.pat_rows.assign(rows, rows + nnz);
    m.pat_cols.assign(cols, cols + nnz);
    build_coloring(m);
}
int orc_n_colors(void* h) { return ((Model*)h)->n_colors; }
int orc_n_state(void* h) { return ((Model*)h)->n_state; }
int orc_n_reduced(void* h) { return ((Model*)h)->n_reduced; }
void orc_set_tprev(void* h, double t) { ((Model*)h)->tprev = t; }
double orc_get_tprev(void* h) { return ((Model*)h)->tprev; }
void orc_set_merged_exact(void* h, int flag) { ((Model*)h)->merged_exact = flag; }
void orc_set_level_prev(void* h, const double* p) {
    Model& m = *(Model*)h; m.level_prev.assign(p, p + m.n_basin);
}

void orc_reduce(void* h, const double* u, double* ur) { ((Model*)h)->reduce_state(u, ur); }

// g(u_reduced, t) -> du ; optionally returns the basin caches
void orc_rhs(void* h, const double* ur, double t, double* du, double* storage, double* level,
             double* area, double* factor) {
    Model& m = *(Model*)h;
    Cache<double> c; c.resize(m);
    water_balance(m, c, ur, t, du);
    for (int i = 0; i < m.n_basin; i++) {
        if (storage) storage[i] = c.storage[i];
        if (level) level[i] = c.level[i];
        if (area) area[i] = c.area[i];
        if (factor) factor[i] = c.factor[i];
    }
}
void orc_jac_dense(void* h, const double* ur, double t, double* J) { jac_dense(*(Model*)h, ur, t, J); }
void orc_jac_colored(void* h, const double* ur, double t, double* vals) {
    jac_colored(*(Model*)h, ur, t, vals);
}
void orc_W_inner(void* h, const double* vals, double gamma, double* W) {
    calc_W_inner(*(Model*)h, vals, gamma, W);
}
int orc_dense_solve(int n, const double* A, const double* b, double* x) {
    std::vector<double> Ac(A, A + (size_t)n * n);
    std::copy(b, b + n, x);
    return dense_solve(n, Ac, x) ? 0 : 1;
}
void orc_limit_flow(void* h, double* u, const double* uprev, double dt, double t) {
    limit_flow(*(Model*)h, u, uprev, dt, t);
}
int orc_step(void* h, double* u, double t, double dt, double abstol, double reltol, int max_iter,
             double* eta_old, int* iters, int* converged, int* negative_storage) {
    StepStats st{};
    int rc = newton_step(*(Model*)h, u, t, dt, abstol, reltol, max_iter, eta_old, &st, true);
    if (iters) *iters = st.iters;
    if (converged) *converged = st.converged;
    if (negative_storage) *negative_storage = st.negative_storage;
    return rc;
}

// DiscreteControl at t = 0 on the state `ur` (RHS flows of that state); out: control states
void orc_apply_discrete_control(void* h, const double* ur, double t) {
    Model& m = *(Model*)h;
    Cache<double> c; c.resize(m);
    std::vector<double> du(m.n_state);
    water_balance(m, c, ur, t, du.data());
    apply_discrete_control(m, c, du.data(), true);
}
void orc_get_dc(void* h, int32_t* state, int32_t* truth, int32_t* counts2) {
    Model& m = *(Model*)h;
    for (int i = 0; i < m.n_dc; i++) state[i] = m.dc_state[i];
    for (size_t i = 0; i < m.dc_truth.size(); i++) truth[i] = m.dc_truth[i];
    counts2[0] = m.dc_changes; counts2[1] = m.dc_undefined;
}

int orc_get_dc_record(void* h, int32_t* out, int max_entries) {
    Model& m = *(Model*)h;
    int n = std::min((int)m.dc_record.size(), max_entries);
    for (int i = 0; i < n; i++) out[i] = m.dc_record[i];
    return (int)m.dc_record.size();
}

// sub-stepping implicit Euler: opts = {abstol, reltol, max_iter, full_newton, grow_iters,
// max_halvings}; out = {iters, substeps, rejects, negative_storage, converged}
int orc_advance(void* h, double* u, double t0, double dt, const double* opts, double* h_try,
                double* eta_old, int* out5) {
    AdvanceOpts o;
    o.abstol = opts[0]; o.reltol = opts[1]; o.max_iter = (int)opts[2]; o.full_newton = (int)opts[3];
    o.grow_iters = (int)opts[4]; o.max_halvings = (int)opts[5]; o.predictor = (int)opts[6];
    o.early_exit = (int)opts[7]; o.shrink_log2 = (int)opts[8];
    StepStats st;
    int rc = advance(*(Model*)h, u, t0, dt, o, h_try, eta_old, &st);
    if (out5) { out5[0] = st.iters; out5[1] = st.substeps; out5[2] = st.rejects;
                out5[3] = st.negative_storage; out5[4] = st.converged; }
    return rc;
}

// scalar helpers for the golden-vector tests
double orc_reduction_factor(double x, double threshold) { return reduction_factor(x, threshold); }
double orc_relaxed_root(double x, double threshold) { return relaxed_root(x, threshold); }
double orc_storage_to_level(void* h, int b, double S) { return storage_to_level(*(Model*)h, b, S); }
double orc_level_to_area(void* h, int b, double level) { return level_to_area(*(Model*)h, b, level); }
double orc_pchip_eval(void* h, int table, double x) { return ((Model*)h)->tabs[table].eval(x); }

// Advance `n` independent instances by `n_steps` outer steps each on `n_threads` threads with
// the same stepping algorithm as the device (CPU baseline: one instance per ensemble member; the
// reference core is single-threaded per model).  opts7 as in orc_advance;
// h_try / eta: per-instance carry; out3 = {failed member-steps, sub-steps, Newton iterations}.
int orc_run_batch(void** hs, double** us, int n, double t0, double dt, int n_steps,
                  const double* opts7, double* h_try, double* eta, int n_threads, int64_t* out3) {
    AdvanceOpts o;
    o.abstol = opts7[0]; o.reltol = opts7[1]; o.max_iter = (int)opts7[2]; o.full_newton = (int)opts7[3];
    o.grow_iters = (int)opts7[4]; o.max_halvings = (int)opts7[5]; o.predictor = (int)opts7[6];
    o.early_exit = (int)opts7[7]; o.shrink_log2 = (int)opts7[8];
    std::vector<std::thread> th;
    std::vector<int64_t> fails(n_threads, 0), subs(n_threads, 0), its(n_threads, 0);
    for (int w = 0; w < n_threads; w++)
        th.emplace_back([&, w]() {
            for (int i = w; i < n; i += n_threads) {
                double t = t0;
                for (int s = 0; s < n_steps; s++) {
                    StepStats st;
                    fails[w] += advance(*(Model*)hs[i], us[i], t, dt, o, &h_try[i], &eta[i], &st);
                    subs[w] += st.substeps;
                    its[w] += st.iters;
                    t += dt;
                }
            }
        });
    for (auto& x : th) x.join();
    if (out3) {
        out3[0] = out3[1] = out3[2] = 0;
        for (int w = 0; w < n_threads; w++) { out3[0] += fails[w]; out3[1] += subs[w]; out3[2] += its[w]; }
    }
    return 0;
}

}  // extern "C"
