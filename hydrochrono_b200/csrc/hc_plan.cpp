// Host-side planning of the radiation look-ahead (no CUDA needed, so the logic is unit-tested on the CPU):
// where every RIRF lag sits on the grid of history rows for a predicted step size, and whether the plan the reference
// would compute for one step (hydro_forces.cpp:343-381,601-610; k_prestep on the device) keeps every lag there.
#include "hc_internal.h"

#include <algorithm>
#include <cmath>

namespace hc {

RadPlan make_rad_plan(const hc_tables& T, double dt, int max_m, int min_lags) {
    RadPlan P;
    const int L = T.L, D = T.D;
    if (L < min_lags || !(dt > 0.0)) return P;
    for (int s = 1; s < L; ++s)
        if (!(T.rirf_t[s] > T.rirf_t[s - 1])) return P;                  // lags must ascend
    if (T.rirf_t[0] < 0.0) return P;
    const double lag_dt = (T.rirf_t.back() - T.rirf_t.front()) / (L - 1);
    const long long m = std::llround(lag_dt / dt);
    P.pnom.assign(L, 0.0);
    if (m >= 1 && m <= max_m && std::fabs(lag_dt - double(m) * dt) <= 1e-6 * dt && T.rirf_t[0] == 0.0) {
        // lag spacing = m dt: lag s sits on history row m s
        P.general = false; P.m = int(m); P.Lk = L;
        for (int s = 0; s < L; ++s) P.pnom[s] = double(m) * s;
    } else {
        // any other ratio: lag s sits between rows floor(x) and floor(x) + 1, x = t_rirf[s] / dt
        for (int s = 0; s < L; ++s) P.pnom[s] = T.rirf_t[s] / dt;
        const double rows = std::floor(P.pnom[L - 1]) + 2.0;
        // FP64 work of the row-grid kernel (rows x D^2) against the HBM traffic of the per-step kernel (2 L rows)
        if (rows < min_lags || rows * D > 64.0 * L) return P;
        P.general = true; P.m = 1; P.Lk = int(rows);
    }
    P.pi.resize(L); P.pw.resize(L);
    for (int s = 0; s < L; ++s) {
        P.pi[s] = int(std::floor(P.pnom[s]));
        P.pw[s] = P.pnom[s] - double(P.pi[s]);
    }
    P.usable = true;
    return P;
}

// tm[0 .. len) = the time history as it will be at the step, newest first (tm[0] = the step's time).  Returns true
// when every lag that has a bracket sits at its nominal position (bracket index + weight of the older row) to within
// `snap` rows; smax = the largest lag with a bracket.  The row-grid kernel needs all lags (full window).
bool rad_plan_step(const hc_tables& T, const RadPlan& P, const double* tm, int len, double snap, int& smax) {
    smax = -1;
    if (!P.usable || len <= 1) return false;
    const int L = T.L;
    const double* rt = T.rirf_t.data();
    const double t0 = tm[0], oldest = tm[len - 1];
    // lags with a bracket: oldest <= t0 - rt[s] (k_prestep's test, same subtraction); rt ascends, so they are 0..lo
    if (!(oldest <= t0 - rt[0])) return false;
    int lo = 0, hi = L - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (oldest <= t0 - rt[mid]) lo = mid; else hi = mid - 1;
    }
    if (P.general && lo != L - 1) return false;
    // every such lag at its nominal position i + w (rows back): |q - (tm[i] - w (tm[i] - tm[i+1]))| <= snap * spacing,
    // i.e. bracket index + older-row weight within snap of the nominal ones (InterpolateVelocity6D's arithmetic)
    const int* pi = P.pi.data();
    const double* pw = P.pw.data();
    for (int s = 0; s <= lo; ++s) {
        const int i = pi[s];
        const double w = pw[s];
        if (i + (w > 0.0 ? 1 : 0) > len - 1) return false;
        const double delta = (i + 1 < len) ? tm[i] - tm[i + 1] : tm[i - 1] - tm[i];
        const double d = (t0 - rt[s]) - (tm[i] - w * delta);
        if (!(std::fabs(d) <= snap * delta)) return false;
    }
    smax = lo;
    return true;
}

// The kernel the block path convolves the history ROWS with, Krow[i][r][c], i = rows back from the step.  Lag grid:
// Krow = (K w) (row m s <-> lag s, stored per lag).  Row grid: the linear interpolation of the velocity between rows
// i and i + 1 at the nominal position x_s = i + wo is folded into the kernel,
//   (K w)[s] (wn v_i + wo v_{i+1})   ->   Krow[i] += wn (K w)[s],   Krow[i + 1] += wo (K w)[s].
dvec rad_plan_row_kernel(const hc_tables& T, const RadPlan& P) {
    const int L = T.L, D = T.D;
    dvec Krow(size_t(P.Lk) * D * D, 0.0);
    for (int s = 0; s < L; ++s) {
        int i = s;
        double wn = 1.0, wo = 0.0;
        if (P.general) {
            i = P.pi[s];
            wo = P.pw[s];
            wn = 1.0 - wo;
        }
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c) {
                const double kw = T.Keff[(size_t(r) * D + c) * L + s] * T.rirf_w[s];
                if (i < P.Lk) Krow[(size_t(i) * D + r) * D + c] += wn * kw;
                if (wo != 0.0 && i + 1 < P.Lk) Krow[(size_t(i + 1) * D + r) * D + c] += wo * kw;
            }
    }
    return Krow;
}

// Next range of work items of a pass that is launched in `nslices` slices: [i0, i1) for the `count` slices after
// `next_slice`.  Nominal boundaries N s / nslices are rounded to multiples of `wave` (whole waves of resident CTAs;
// 1 = equal slices).  The range always starts where the previous one ended, whatever `wave` was then, and the last
// slice always ends at N, so the slices of a pass cover [0, N) exactly once even if the policy changes on the way.
bool rad_pass_next(long long N, int nslices, int& next_slice, long long& next_item, int count, long long wave,
                   long long& i0, long long& i1) {
    const int s1 = std::min(nslices, next_slice + std::max(count, 0));
    if (wave < 1) wave = 1;
    long long end = N;
    if (s1 < nslices) {
        const long long x = N * s1 / nslices;
        end = std::min(N, (x + wave / 2) / wave * wave);
    }
    i0 = next_item;
    i1 = std::max(i0, end);
    next_slice = s1;
    next_item = i1;
    return i1 > i0;
}

}  // namespace hc

extern "C" {

hc_status hc_rad_lookahead_plan(const hc_tables* t, double dt_hint, int* mode, int* rows_per_lag, int* kernel_lags) {
    if (!t) { hc::set_last_error("null argument"); return HC_ERR_INVALID; }
    const hc::RadPlan P = hc::make_rad_plan(*t, dt_hint, 8, 16);
    if (mode) *mode = !P.usable ? 0 : (P.general ? 2 : 1);
    if (rows_per_lag) *rows_per_lag = P.usable ? P.m : 0;
    if (kernel_lags) *kernel_lags = P.usable ? P.Lk : 0;
    return HC_OK;
}

int hc_rad_pass_next(long long items, int nslices, int* next_slice, long long* next_item, int count, long long wave,
                     long long* i0, long long* i1) {
    return hc::rad_pass_next(items, nslices, *next_slice, *next_item, count, wave, *i0, *i1) ? 1 : 0;
}

hc_status hc_rad_lookahead_row_kernel(const hc_tables* t, double dt_hint, double* out) {
    if (!t || !out) { hc::set_last_error("null argument"); return HC_ERR_INVALID; }
    const hc::RadPlan P = hc::make_rad_plan(*t, dt_hint, 8, 16);
    if (!P.usable) { hc::set_last_error("the radiation look-ahead cannot serve this step size"); return HC_ERR_INVALID; }
    const hc::dvec K = hc::rad_plan_row_kernel(*t, P);
    std::copy(K.begin(), K.end(), out);
    return HC_OK;
}

hc_status hc_rad_lookahead_check_step(const hc_tables* t, double dt_hint, double bracket_snap, const double* times_newest_first,
                                      int n, int* smax) {
    if (!t || !times_newest_first || !smax) { hc::set_last_error("null argument"); return HC_ERR_INVALID; }
    const hc::RadPlan P = hc::make_rad_plan(*t, dt_hint, 8, 16);
    int sm = -1;
    const bool ok = hc::rad_plan_step(*t, P, times_newest_first, n, bracket_snap, sm);
    *smax = ok ? sm : -1;
    return HC_OK;
}

}  // extern "C"
