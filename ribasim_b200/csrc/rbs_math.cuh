// rbs_math.cuh — scalar fp64 device functions of the RHS and their analytic partials.
//
// Value paths keep the reference's operation order (no FMA contraction: the library is built
// with -fmad=false) so that results are reproducible against the CPU restatement; the partial
// derivatives are what ForwardDiff produces from the same expressions, written out by hand.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define RBS_INF (__longlong_as_double(0x7ff0000000000000LL))

// reduction_factor — core/src/util.jl:312-321.  Returns phi, writes dphi/dx.
__device__ __forceinline__ double rbs_reduction_factor(double x, double thr, double& dphi) {
    if (x < 0.0) { dphi = 0.0; return 0.0; }
    if (x < thr) {
        double xs = x / thr;
        dphi = 6.0 * xs * (1.0 - xs) / thr;
        return (-2.0 * xs + 3.0) * (xs * xs);
    }
    dphi = 0.0;
    return 1.0;
}
__device__ __forceinline__ double rbs_reduction_factor(double x, double thr) {
    if (x < 0.0) return 0.0;
    if (x < thr) {
        double xs = x / thr;
        return (-2.0 * xs + 3.0) * (xs * xs);
    }
    return 1.0;
}

// relaxed_root — core/src/util.jl:679-685.  Writes d/dx and d/dthreshold.
__device__ __forceinline__ double rbs_relaxed_root(double x, double thr, double& dx, double& dthr) {
    if (fabs(x) < thr) {
        double r = x / thr;
        double st = sqrt(thr);
        double xs = x / st;
        dx = 0.25 * (5.0 - 3.0 * (r * r)) / st;
        dthr = 0.625 * (xs / thr) * (r * r - 1.0);
        return 0.25 * xs * (5.0 - r * r);
    }
    double s = sqrt(fabs(x));
    dx = 0.5 / s;
    dthr = 0.0;
    return x > 0.0 ? s : (x < 0.0 ? -s : 0.0 * s);
}

// Julia's max(x, y): NaN propagates (CUDA fmax would drop it).
__device__ __forceinline__ double rbs_jl_max(double a, double b) {
    if (a != a) return a;
    return a > b ? a : b;
}

// Basin profile: storage -> level (LinearInterpolationIntInv, Linear extrapolation both
// sides; core/src/read.jl:2281-2296, evaluated at core/src/util.jl:42-44) and
// level -> area (LinearInterpolation, Constant extrapolation; read.jl:2303-2315).
struct RbsBasinEval {
    double level, dlevel;  // h(S), dh/dS
    double area, darea;    // A(h(S)), dA/dS
};

__device__ __forceinline__ RbsBasinEval rbs_basin_eval(double S, int n, const double* __restrict__ h,
                                                       const double* __restrict__ a,
                                                       const double* __restrict__ a_slope,
                                                       const double* __restrict__ cs,
                                                       const double* __restrict__ area,
                                                       const double* __restrict__ area_slope) {
    RbsBasinEval r;
    if (S < cs[0]) {
        r.level = h[0] + (S - cs[0]) / a[0];
        r.dlevel = 1.0 / a[0];
    } else if (S > cs[n - 1]) {
        r.level = h[n - 1] + (S - cs[n - 1]) / a[n - 1];
        r.dlevel = 1.0 / a[n - 1];
    } else {
        int i = 0;  // searchsortedlast(cs, S) clamped to [1, n-1], 0-based
        while (i + 2 < n && S >= cs[i + 1]) i++;
        double dS = S - cs[i];
        double x = a[i];
        double m = a_slope[i];
        // cancellation-free quadratic root; dh/dS = 1/root
        double root = sqrt(x * x + m * (2.0 * dS));
        r.level = h[i] + 2.0 * dS / (x + root);
        r.dlevel = 1.0 / root;
    }
    double lv = r.level;
    if (lv < h[0]) {
        r.area = area[0];
        r.darea = 0.0;
    } else if (lv > h[n - 1]) {
        r.area = area[n - 1];
        r.darea = 0.0;
    } else {
        int i = 0;
        while (i + 2 < n && lv >= h[i + 1]) i++;
        r.area = area[i] + area_slope[i] * (lv - h[i]);
        r.darea = area_slope[i] * r.dlevel;
    }
    return r;
}

// CubicHermiteSpline evaluation with the table's extrapolation slopes
// (DataInterpolations; call sites core/src/solve.jl:469, :109).  Writes dQ/dx.
__device__ __forceinline__ double rbs_pchip_eval(double x, int n, const double* __restrict__ t,
                                                 const double* __restrict__ u,
                                                 const double* __restrict__ du,
                                                 const double* __restrict__ c1,
                                                 const double* __restrict__ c2, double slope_left,
                                                 double slope_right, double& deriv) {
    if (x < t[0]) {
        deriv = slope_left;
        return u[0] + slope_left * (x - t[0]);
    }
    if (x > t[n - 1]) {
        deriv = slope_right;
        return u[n - 1] + slope_right * (x - t[n - 1]);
    }
    int i = 0;
    while (i + 2 < n && x >= t[i + 1]) i++;
    double d0 = x - t[i];
    double d1 = x - t[i + 1];
    double out = u[i] + d0 * du[i];
    out += (d0 * d0) * (c1[i] + d1 * c2[i]);
    deriv = du[i] + 2.0 * d0 * (c1[i] + d1 * c2[i]) + (d0 * d0) * c2[i];
    return out;
}

// linear_resistance_flow — core/src/solve.jl:431-450
__device__ __forceinline__ double rbs_lr_flow(double h_a, double h_b, double f_a, double f_b,
                                              double dh_a, double dh_b, double df_a, double df_b,
                                              double resistance, double qmax, double& dq_a,
                                              double& dq_b) {
    double dh = h_a - h_b;
    double q_unlimited = dh / resistance;
    bool hi = q_unlimited > qmax, lo = q_unlimited < -qmax;
    double q = hi ? qmax : (lo ? -qmax : q_unlimited);
    double g = (hi || lo) ? 0.0 : 1.0 / resistance;
    bool up = q_unlimited > 0.0;
    double f = up ? f_a : f_b;
    dq_a = g * dh_a * f + (up ? q * df_a : 0.0);
    dq_b = -g * dh_b * f + (up ? 0.0 : q * df_b);
    return q * f;
}

// manning_resistance_flow — core/src/solve.jl:528-589
__device__ __forceinline__ double rbs_manning_flow(double h_a, double h_b, double f_a, double f_b,
                                                   double dh_a, double dh_b, double df_a,
                                                   double df_b, double L, double n, double width,
                                                   double slope, double bottom_a, double bottom_b,
                                                   double& dq_a, double& dq_b, bool want_jac) {
    double d_a = h_a - bottom_a;
    double d_b = h_b - bottom_b;
    double d = 0.5 * (d_a + d_b);
    double A_a = width * d + slope * (d_a * d_a);
    double A_b = width * d + slope * (d_b * d_b);
    double A = 0.5 * (A_a + A_b);
    double sul = sqrt(slope * slope + 1.0);
    double P_a = width + 2.0 * d_a * sul;
    double P_b = width + 2.0 * d_b * sul;
    double R_a = A_a / P_a;
    double R_b = A_b / P_b;
    double R = 0.5 * (R_a + R_b);
    double dh = h_a - h_b;
    const double nu = 1.004e-6;
    double cb = cbrt(R);
    double T = (2000 * nu) * n * cb / A;
    double thr_raw = T * T;
    double thr = rbs_jl_max(thr_raw, 1.0e-5);
    double g = cbrt(R * R);
    double rr_dx, rr_dthr;
    double x = dh / L;
    double rr = rbs_relaxed_root(x, thr, rr_dx, rr_dthr);
    double An = A / n;
    double q0 = An * g * rr;
    bool up = q0 > 0.0;
    double f = up ? f_a : f_b;
    if (want_jac) {
        if (A == 0.0 || R == 0.0) {  // 0 * inf in the formal derivative: exactly dry reach
            dq_a = 0.0;
            dq_b = 0.0;
        } else {
            // derivatives with respect to h_a and h_b
            double dAa_a = 0.5 * width + 2.0 * slope * d_a, dAa_b = 0.5 * width;
            double dAb_a = 0.5 * width, dAb_b = 0.5 * width + 2.0 * slope * d_b;
            double dA_a = 0.5 * (dAa_a + dAb_a), dA_b = 0.5 * (dAa_b + dAb_b);
            double dRa_a = (dAa_a * P_a - A_a * (2.0 * sul)) / (P_a * P_a);
            double dRa_b = dAa_b / P_a;
            double dRb_a = dAb_a / P_b;
            double dRb_b = (dAb_b * P_b - A_b * (2.0 * sul)) / (P_b * P_b);
            double dR_a = 0.5 * (dRa_a + dRb_a), dR_b = 0.5 * (dRa_b + dRb_b);
            double inv3cb2 = 1.0 / (3.0 * (cb * cb));
            double c = (2000 * nu) * n;
            bool thr_active = thr_raw > 1.0e-5;
            double dT_a = c * (dR_a * inv3cb2 * A - cb * dA_a) / (A * A);
            double dT_b = c * (dR_b * inv3cb2 * A - cb * dA_b) / (A * A);
            double dthr_a = thr_active ? 2.0 * T * dT_a : 0.0;
            double dthr_b = thr_active ? 2.0 * T * dT_b : 0.0;
            double dg_a = 2.0 * R * dR_a / (3.0 * (g * g));
            double dg_b = 2.0 * R * dR_b / (3.0 * (g * g));
            double dq0_a = (dA_a / n) * g * rr + An * dg_a * rr +
                           An * g * (rr_dx * (1.0 / L) + rr_dthr * dthr_a);
            double dq0_b = (dA_b / n) * g * rr + An * dg_b * rr +
                           An * g * (rr_dx * (-1.0 / L) + rr_dthr * dthr_b);
            dq_a = dq0_a * dh_a * f + (up ? q0 * df_a : 0.0);
            dq_b = dq0_b * dh_b * f + (up ? 0.0 : q0 * df_b);
        }
    }
    return q0 * f;
}

// tabulated_rating_curve_flow — core/src/solve.jl:452-474
__device__ __forceinline__ double rbs_trc_flow(double h_a, double h_b, double f_a, double dh_a,
                                               double dh_b, double df_a, double Q, double dQ,
                                               double max_down, double ldt, double& dq_a,
                                               double& dq_b) {
    double dh = h_a - h_b;
    double r1p, r2p;
    double r1 = rbs_reduction_factor(dh, ldt, r1p);
    double r2 = rbs_reduction_factor(max_down - h_b, ldt, r2p);
    double q0 = f_a * Q;
    double q1 = q0 * r1;
    double q = q1 * r2;
    dq_a = ((df_a * Q + f_a * dQ * dh_a) * r1 + q0 * r1p * dh_a) * r2;
    dq_b = (q0 * (r1p * (-dh_b))) * r2 + q1 * (r2p * (-dh_b));
    return q;
}

// formulate_pump_or_outlet_flow! — core/src/solve.jl:657-770 (one node).
// g_fr = dq/dflow_rate, dq_a / dq_b = direct partials wrt the source / destination storages.
__device__ __forceinline__ double rbs_pump_outlet_flow(double flow_rate, double h_a, double h_b,
                                                       double f_a, double dh_a, double dh_b,
                                                       double df_a, double lower, double upper,
                                                       double flow_demand, double min_up,
                                                       double max_down, double ldt, bool outlet,
                                                       double& g_fr, double& dq_a, double& dq_b) {
    double q1 = flow_rate * f_a;
    if (flow_demand == flow_demand)  // FlowDemand as a lower bound (solve.jl:719-740)
        lower = flow_demand > upper ? upper : (flow_demand < lower ? lower : flow_demand);
    bool hi = q1 > upper, lo = q1 < lower;
    double q2 = hi ? upper : (lo ? lower : q1);
    double cl = (hi || lo) ? 0.0 : 1.0;
    double rd = 1.0, rdp = 0.0;
    double q = q2;
    if (outlet) {
        rd = rbs_reduction_factor(h_a - h_b, ldt, rdp);
        q *= rd;
    }
    double rup, rdnp;
    double ru = rbs_reduction_factor(h_a - min_up, ldt, rup);
    q *= ru;
    double rdn = rbs_reduction_factor(max_down - h_b, ldt, rdnp);
    q *= rdn;
    double R = rd * ru * rdn;
    g_fr = cl * f_a * R;
    dq_a = cl * flow_rate * df_a * R + q2 * ((rdp * dh_a) * ru * rdn + rd * (rup * dh_a) * rdn);
    dq_b = q2 * ((rdp * (-dh_b)) * ru * rdn + rd * ru * (rdnp * (-dh_b)));
    return q;
}
