// Setup-time wave mathematics on the host: frequency grids, spectra, dispersion, random phases and the
// excitation-IRF resampling.  Reference: src/wave_types.cpp:178-255 (ComputeWaveNumber), :572-628
// (ResampleIRF, GetWidthArray), :643-715 (CreateSpectrum, PM / JONSWAP).  The heavy part of the irregular
// set-up -- eta synthesis for every realisation -- runs on the device (hc_kernels.cu).
#include <algorithm>
#include <random>

#include "hc_internal.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace hc {

// Eigen::VectorXd::LinSpaced(n, lo, hi) (Eigen 3.4, floating scalars): lo + i*step, end point written exactly;
// when |hi| < |lo| Eigen counts down from hi instead.
dvec linspaced(int n, double lo, double hi) {
    dvec v(n > 0 ? n : 0);
    if (n <= 0) return v;
    if (n == 1) { v[0] = hi; return v; }
    const int last = n - 1;
    const double step = (hi - lo) / double(last);
    if (std::fabs(hi) < std::fabs(lo)) {
        v[0] = lo;
        for (int i = 1; i < n; ++i) v[i] = hi - double(last - i) * step;
    } else {
        for (int i = 0; i < last; ++i) v[i] = lo + double(i) * step;
        v[last] = hi;
    }
    return v;
}

// Newton iteration on omega^2 = g k tanh(k h); deep-water shortcut for h == 0, h > 1000 or inf.
double wave_number(double omega, double depth, double g) {
    const double tol = 1e-6;
    const int max_it = 100;
    if (omega <= 0.0) fail(HC_ERR_INVALID, "Angular frequency must be positive.");
    if (depth < 0.0) fail(HC_ERR_INVALID, "Water depth cannot be negative.");
    if (g <= 0.0) fail(HC_ERR_INVALID, "Gravity must be positive.");
    const double k_deep = omega * omega / g;
    if (depth == 0.0 || depth > 1000.0 || std::isinf(depth)) return k_deep;
    double k = k_deep, err = 1.0;
    int it = 0;
    for (; err > tol && it < max_it; ++it) {
        const double th = std::tanh(k * depth);
        const double fk = omega * omega - g * k * th;
        const double dfk = -2.0 * g * th - g * k * depth * (1.0 - th * th);
        if (std::fabs(dfk) < tol) fail(HC_ERR_INVALID, "Numerical instability: derivative too close to zero.");
        const double dk = fk / dfk;
        k -= dk;
        err = std::fabs(dk);
    }
    if (it >= max_it) fail(HC_ERR_INVALID, "Failed to converge within maximum iterations.");
    return k;
}

dvec pierson_moskowitz(dvec f, double Hs, double Tp) {
    std::sort(f.begin(), f.end());
    dvec S(f.size());
    for (size_t i = 0; i < f.size(); ++i)
        S[i] = 1.25 * std::pow(1 / Tp, 4) * std::pow(Hs / 2, 2) * std::pow(f[i], -5) *
               std::exp(-1.25 * std::pow(1 / Tp, 4) * std::pow(f[i], -4));
    return S;
}

dvec jonswap(dvec f, double Hs, double Tp, double gamma, bool normalized) {
    std::sort(f.begin(), f.end());
    dvec S = pierson_moskowitz(f, Hs, Tp);
    const double norm = (1 - 0.287 * std::log(gamma));
    for (size_t i = 0; i < S.size(); ++i) {
        const double sigma = f[i] <= 1.0 / Tp ? 0.07 : 0.09;
        S[i] *= std::pow(gamma, std::exp(-(1.0 / (2.0 * std::pow(sigma, 2))) * std::pow(f[i] * Tp - 1.0, 2)));
        if (normalized) S[i] *= norm;
    }
    return S;
}

// std::uniform_real_distribution<double>(0, 2pi) over std::mt19937(seed): generate_canonical<double,53> takes two
// 32-bit draws, low word first.  Spelled out so the stream does not depend on the C++ library in use.
dvec random_phases(int seed, int n) {
    std::mt19937 gen(static_cast<std::mt19937::result_type>(seed));
    const double two32 = 4294967296.0;
    dvec out(n);
    for (int i = 0; i < n; ++i) {
        const double lo = double(gen());
        const double hi = double(gen());
        double u = (lo + hi * two32) / (two32 * two32);
        if (u >= 1.0) u = std::nextafter(1.0, 0.0);
        out[i] = u * (2 * M_PI - 0.0) + 0.0;
    }
    return out;
}

// ---- cubic B-spline interpolation with averaged knots (Eigen unsupported/Splines) ---------------------------
namespace {
struct CubicBasis {
    dvec knots;
    int n;
    explicit CubicBasis(int n_) : knots(n_ + 4), n(n_) {
        // parameters u_i = LinSpaced(n, 0, 1); KnotAveraging: interior knot j+3 = mean(u_j, u_j+1, u_j+2)
        dvec u = linspaced(n, 0.0, 1.0);
        for (int j = 1; j < n - 3; ++j) knots[j + 3] = (u[j] + u[j + 1] + u[j + 2]) / 3.0;
        for (int j = 0; j < 4; ++j) { knots[j] = 0.0; knots[n + j] = 1.0; }
    }
    int span(double x) const {
        if (x <= knots[0]) return 3;
        const double* first = knots.data() + 2;
        const double* last = knots.data() + knots.size() - 4;
        return int(std::upper_bound(first, last, x) - knots.data()) - 1;
    }
    // the four non-zero cubic basis functions on `sp` (Cox-de Boor triangle)
    void eval(double x, int sp, double N[4]) const {
        double dl[4], dr[4];
        N[0] = 1.0;
        for (int j = 1; j <= 3; ++j) {
            dl[j] = x - knots[sp + 1 - j];
            dr[j] = knots[sp + j] - x;
            double carry = 0.0;
            for (int r = 0; r < j; ++r) {
                const double q = N[r] / (dr[r + 1] + dl[j - r]);
                N[r] = carry + dr[r + 1] * q;
                carry = dl[j - r] * q;
            }
            N[j] = carry;
        }
    }
};
}  // namespace

void bspline_resample(const double* pts, int dim, int n, int m, double* out) {
    if (n < 4) fail(HC_ERR_INVALID, "excitation IRF needs at least 4 samples for cubic resampling");
    CubicBasis B(n);
    dvec u = linspaced(n, 0.0, 1.0);
    // Collocation matrix: row i has <= 4 entries starting at column span-3; rows 0 and n-1 are unit rows.
    // Banded Gaussian elimination with row pivoting inside the band (half-bandwidth 3 below / up to 6 above).
    const int kl = 3, ku = 6, W = kl + ku + 1;
    std::vector<double> A(size_t(n) * W, 0.0);       // A(i, j) stored at A[i*W + (j - i + kl)]
    auto at = [&](int i, int j) -> double& { return A[size_t(i) * W + (j - i + kl)]; };
    at(0, 0) = 1.0;
    at(n - 1, n - 1) = 1.0;
    for (int i = 1; i < n - 1; ++i) {
        const int sp = B.span(u[i]);
        double Nb[4];
        B.eval(u[i], sp, Nb);
        for (int q = 0; q < 4; ++q) at(i, sp - 3 + q) = Nb[q];
    }
    std::vector<double> X(size_t(n) * dim);
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d) X[size_t(i) * dim + d] = pts[size_t(d) * n + i];
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = std::fabs(at(k, k));
        const int rlim = std::min(n - 1, k + kl);
        for (int r = k + 1; r <= rlim; ++r)
            if (std::fabs(at(r, k)) > best) { best = std::fabs(at(r, k)); piv = r; }
        if (best == 0.0) fail(HC_ERR_INVALID, "singular spline collocation matrix");
        const int clim = std::min(n - 1, k + ku);
        if (piv != k) {
            for (int c = k; c <= clim; ++c) std::swap(at(k, c), at(piv, c));
            for (int d = 0; d < dim; ++d) std::swap(X[size_t(k) * dim + d], X[size_t(piv) * dim + d]);
        }
        for (int r = k + 1; r <= rlim; ++r) {
            const double mlt = at(r, k) / at(k, k);
            if (mlt == 0.0) continue;
            for (int c = k; c <= clim; ++c) at(r, c) -= mlt * at(k, c);
            for (int d = 0; d < dim; ++d) X[size_t(r) * dim + d] -= mlt * X[size_t(k) * dim + d];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        const int clim = std::min(n - 1, k + ku);
        for (int d = 0; d < dim; ++d) {
            double s = X[size_t(k) * dim + d];
            for (int c = k + 1; c <= clim; ++c) s -= at(k, c) * X[size_t(c) * dim + d];
            X[size_t(k) * dim + d] = s / at(k, k);
        }
    }
    dvec v = linspaced(m, 0.0, 1.0);
    for (int j = 0; j < m; ++j) {
        const int sp = B.span(v[j]);
        double Nb[4];
        B.eval(v[j], sp, Nb);
        for (int d = 0; d < dim; ++d) {
            double s = 0.0;
            for (int q = 0; q < 4; ++q) s += Nb[q] * X[size_t(sp - 3 + q) * dim + d];
            out[size_t(d) * m + j] = s;
        }
    }
}

// IrregularWaves::InitializeIRFVectors + ResampleIRF + CalculateWidthIRF
std::vector<ExcIrfBody> resample_excitation_irf(const hc_tables& T, double dt) {
    if (T.Le0 <= 0) fail(HC_ERR_INVALID, "tables hold no excitation impulse response data");
    std::vector<ExcIrfBody> out(T.N);
    for (int b = 0; b < T.N; ++b) {
        const BodyTables& B = T.body[b];
        ExcIrfBody& E = out[b];
        if (dt > 0.0) {
            const double t0 = B.exc_irf_t.front(), t1 = B.exc_irf_t.back();
            const int m = static_cast<int>(std::ceil((t1 - t0) / dt));
            if (m < 2) fail(HC_ERR_INVALID, "simulation_dt too large for the excitation IRF window");
            E.t = linspaced(m, t0, t1);
            E.f.resize(size_t(6) * m);
            bspline_resample(B.exc_irf_f.data(), 6, T.Le0, m, E.f.data());
        } else {
            E.t = B.exc_irf_t;
            E.f = B.exc_irf_f;
        }
        E.w = trapezoid_widths(E.t);
    }
    return out;
}

}  // namespace hc

extern "C" {

hc_status hc_pierson_moskowitz_spectrum_hz(int n, const double* f, double Hs, double Tp, double* S) {
    try {
        hc::dvec s = hc::pierson_moskowitz(hc::dvec(f, f + n), Hs, Tp);
        std::copy(s.begin(), s.end(), S);
        return HC_OK;
    } catch (const std::exception& e) { hc::set_last_error(e.what()); return HC_ERR_INVALID; }
}
hc_status hc_jonswap_spectrum_hz(int n, const double* f, double Hs, double Tp, double gamma, int is_normalized,
                                 double* S) {
    try {
        hc::dvec s = hc::jonswap(hc::dvec(f, f + n), Hs, Tp, gamma, is_normalized != 0);
        std::copy(s.begin(), s.end(), S);
        return HC_OK;
    } catch (const std::exception& e) { hc::set_last_error(e.what()); return HC_ERR_INVALID; }
}
hc_status hc_compute_wave_number(double omega, double water_depth, double g, double* k) {
    try {
        *k = hc::wave_number(omega, water_depth, g);
        return HC_OK;
    } catch (const hc::StatusError& e) { hc::set_last_error(e.msg); return e.code; }
}

hc_status hc_resample_excitation_irf(const hc_tables* t, double dt, int body, int* n_out, double* t_out,
                                     double* width_out, double* f_out) {
    try {
        if (!t || body < 0 || body >= t->N) { hc::set_last_error("body index out of range"); return HC_ERR_OUT_OF_RANGE; }
        std::vector<hc::ExcIrfBody> all = hc::resample_excitation_irf(*t, dt);
        const hc::ExcIrfBody& E = all[body];
        if (n_out) *n_out = int(E.t.size());
        if (t_out) std::copy(E.t.begin(), E.t.end(), t_out);
        if (width_out) std::copy(E.w.begin(), E.w.end(), width_out);
        if (f_out) std::copy(E.f.begin(), E.f.end(), f_out);
        return HC_OK;
    } catch (const hc::StatusError& e) { hc::set_last_error(e.msg); return e.code; }
      catch (const std::exception& e) { hc::set_last_error(e.what()); return HC_ERR_INVALID; }
}
// Airy kinematics (src/wave_types.cpp:14-160,515-545); operation order of the reference, component by component.
hc_status hc_wave_kinematics(int n, const double* omega, const double* amplitude, const double* phase,
                             const double* wavenumber, const double position[3], double time, double water_depth,
                             double mwl, int wheeler_stretching, double* eta_out, double velocity[3],
                             double acceleration[3]) {
    if (n < 0 || (n > 0 && (!omega || !amplitude || !phase || !wavenumber)) || !position) {
        hc::set_last_error("null argument");
        return HC_ERR_INVALID;
    }
    const double x_pos = position[0];
    double eta = 0.0;
    for (int i = 0; i < n; ++i) eta += amplitude[i] * std::cos(wavenumber[i] * x_pos - omega[i] * time + phase[i]);
    if (eta_out) *eta_out = eta;
    double pz = position[2];
    if (wheeler_stretching) {
        const double z_rel = position[2] - mwl;
        pz = water_depth * (z_rel - eta) / (water_depth + eta);    // replaces position.z(); mwl is subtracted again below
    }
    const double z_pos = pz - mwl;
    double v[3] = {0.0, 0.0, 0.0}, a[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n; ++i) {
        const double k = wavenumber[i], om = omega[i], A = amplitude[i];
        const double th = k * x_pos - om * time + phase[i];
        if (2 * M_PI / k > water_depth || k * water_depth > 500.0) {        // deep water
            v[0] += om * A * std::exp(k * z_pos) * std::cos(th);
            v[2] += om * A * std::exp(k * z_pos) * std::sin(th);
            a[0] += om * om * A * std::exp(k * z_pos) * std::sin(th);
            a[2] += -om * om * A * std::exp(k * z_pos) * std::cos(th);
        } else {                                                            // finite depth
            v[0] += om * A * std::cosh(k * (z_pos + water_depth)) / std::sinh(k * water_depth) * std::cos(th);
            v[2] += om * A * std::sinh(k * (z_pos + water_depth)) / std::sinh(k * water_depth) * std::sin(th);
            a[0] += om * om * A * std::cosh(k * (z_pos + water_depth)) / std::sinh(k * water_depth) * std::sin(th);
            a[2] += -om * om * A * std::sinh(k * (z_pos + water_depth)) / std::sinh(k * water_depth) * std::cos(th);
        }
    }
    if (velocity) std::copy(v, v + 3, velocity);
    if (acceleration) std::copy(a, a + 3, acceleration);
    return HC_OK;
}

hc_status hc_random_phases(int seed, int n, double* out) {
    hc::dvec v = hc::random_phases(seed, n);
    std::copy(v.begin(), v.end(), out);
    return HC_OK;
}

}  // extern "C"
