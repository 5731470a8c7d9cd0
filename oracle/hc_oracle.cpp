// =====================================================================================
// hc_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A line-faithful C++17 restatement of HydroChrono's per-timestep hydrodynamic force path,
// free of Chrono/Eigen/HDF5 (none are available in this image, so the reference itself cannot
// be compiled; see DESIGN.md "Oracle").  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product
// (hydrochrono_b200/csrc) never includes, links or calls anything in this directory.
//
// Pinning: validated against the reference's own golden trajectories for the sphere
// (decay, regular waves #1..#10, irregular waves) in tests/test_oracle_goldens.py.
// Force-level parity is "unpinned" by the reference (it ships no force-level vectors,
// SURVEY.md F6); the trajectory goldens pin every function below at 1e-6 print precision.
//
// Every function cites the reference file:line (relative to /root/reference) it restates.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace orc {

constexpr int kDofPerBody = 6;   // src/hydro_forces.cpp:35
constexpr int kDofLinOrRot = 3;  // src/hydro_forces.cpp:36

using Vec = std::vector<double>;

// Eigen::VectorXd::LinSpaced(n, lo, hi) for floating scalars (Eigen 3.4 linspaced_op_impl):
// step=(hi-lo)/(n-1); flip when |hi|<|lo|; last (or first, when flipped) element exact.
static Vec LinSpaced(int n, double lo, double hi) {
    Vec out(std::max(n, 0));
    if (n <= 0) return out;
    if (n == 1) { out[0] = hi; return out; }  // Eigen: size 1 -> high
    const int size1 = n - 1;
    const double step = (hi - lo) / double(size1);
    const bool flip = std::abs(hi) < std::abs(lo);
    for (int i = 0; i < n; ++i) {
        if (flip) out[i] = (i == 0) ? lo : (hi - double(size1 - i) * step);
        else      out[i] = (i == size1) ? hi : (lo + double(i) * step);
    }
    return out;
}

// src/wave_types.cpp:608-620 GetWidthArray, and src/hydro_forces.cpp:181-190 (same formula).
static Vec WidthArray(const Vec& x) {
    Vec w(x.size());
    const int n = (int)x.size();
    for (int ii = 0; ii < n; ii++) {
        w[ii] = 0.0;
        if (ii < n - 1) w[ii] += 0.5 * std::abs(x[ii + 1] - x[ii]);
        if (ii > 0)     w[ii] += 0.5 * std::abs(x[ii] - x[ii - 1]);
    }
    return w;
}

// ------------------------------------------------------------------------------------
// Tables: HydroData after H5FileInfo::ReadH5Data (src/h5fileinfo.cpp:27-91) with the scalings
// applied where the reference applies them.
// ------------------------------------------------------------------------------------
struct Tables {
    int N = 0, D = 0, L = 0, nw = 0, Le0 = 0;
    double rho = 0, g = 0, depth = 0;
    Vec rirf_t;                  // [L]  (body 0; others checked equal to 1e-10, h5fileinfo.cpp:329-343)
    Vec rirf_w;                  // [L]  hydro_forces.cpp:181-190
    Vec K;                       // [N][6][D][L] RAW file values; rho applied on access (h5fileinfo.cpp:321-323)
    Vec Kproc;                   // TaperedDirect processed kernel (already x rho) or empty
    bool tapered = false;
    Vec lin;                     // [N][6][6] unscaled
    Vec ainf;                    // [N][6][D] x rho (h5fileinfo.cpp:60-61)
    Vec disp_vol;                // [N]
    Vec cg, cb;                  // [N][3]
    Vec equilibrium;             // [D]   hydro_forces.cpp:208-216
    Vec cb_minus_cg;             // [3N]
    Vec w_list;                  // [nw]
    Vec exc_mag;                 // [N][6][nw] x rho*g (h5fileinfo.cpp:73-75)
    Vec exc_phase;               // [N][6][nw]
    Vec exc_irf_t;               // [N][Le0]
    Vec exc_irf_f;               // [N][6][Le0] x rho*g (h5fileinfo.cpp:90)

    // HydroData::GetRIRFVal (h5fileinfo.cpp:321-323) via TestHydro::GetRIRFval (hydro_forces.cpp:693-711)
    inline double RIRFval(int row, int col, int st) const {
        const int body = row / kDofPerBody, row_dof = row % kDofPerBody;
        const size_t idx = ((size_t(body) * 6 + row_dof) * D + col) * L + st;
        if (tapered) return Kproc[idx];
        return K[idx] * rho;
    }
};

struct TaperedOpts {  // include/hydroc/hydro_forces.h:246-259
    int smoothing_moving_average = 0;  // 0: SG 5-point (default branch), 1: moving_average
    int window_length = 5;
    double rirf_end_time = -1.0;
    double taper_start_percent = 0.8;
    double taper_end_percent = 1.0;
    double taper_final_amplitude = 0.0;
};

// src/hydro_forces.cpp:385-535 EnsureProcessedRIRF
static void ProcessRIRF(Tables& T, const TaperedOpts& o) {
    const int steps = T.L, cols = T.D, rows = kDofPerBody;
    T.Kproc.assign(T.K.size(), 0.0);
    const double sg5[5] = {-3.0 / 35.0, 12.0 / 35.0, 17.0 / 35.0, 12.0 / 35.0, -3.0 / 35.0};
    for (int b = 0; b < T.N; ++b) {
        int effective_steps = steps;
        if (o.rirf_end_time > 0.0) {
            double dt = T.rirf_t[1] - T.rirf_t[0];
            int end_step = static_cast<int>(std::floor(o.rirf_end_time / dt));
            effective_steps = std::min(end_step, steps);
        }
        for (int row_dof = 0; row_dof < rows; ++row_dof) {
            for (int col = 0; col < cols; ++col) {
                const size_t base = ((size_t(b) * 6 + row_dof) * T.D + col) * T.L;
                Vec k_raw(steps);
                for (int s = 0; s < steps; ++s) k_raw[s] = T.K[base + s] * T.rho;
                if (o.rirf_end_time > 0.0) k_raw.resize(effective_steps);
                Vec k_smooth(effective_steps);
                if (o.smoothing_moving_average) {
                    const int w = std::max(3, o.window_length);
                    const int half = w / 2;
                    for (int s = 0; s < effective_steps; ++s) {
                        int a = std::max(0, s - half);
                        int bb = std::min(effective_steps - 1, s + half);
                        double sum = 0.0; int cnt = 0;
                        for (int i = a; i <= bb; ++i) { sum += k_raw[i]; ++cnt; }
                        k_smooth[s] = (cnt > 0) ? (sum / cnt) : k_raw[s];
                    }
                } else {
                    if (effective_steps >= 5) {
                        k_smooth[0] = k_raw[0];
                        k_smooth[1] = k_raw[1];
                        for (int s = 2; s <= effective_steps - 3; ++s) {
                            k_smooth[s] = sg5[0] * k_raw[s - 2] + sg5[1] * k_raw[s - 1] + sg5[2] * k_raw[s] +
                                          sg5[3] * k_raw[s + 1] + sg5[4] * k_raw[s + 2];
                        }
                        k_smooth[effective_steps - 2] = k_raw[effective_steps - 2];
                        k_smooth[effective_steps - 1] = k_raw[effective_steps - 1];
                    } else {
                        k_smooth = k_raw;
                    }
                }
                int tc_index = static_cast<int>(std::floor(o.taper_start_percent * static_cast<double>(effective_steps)));
                int tc_end = static_cast<int>(std::floor(o.taper_end_percent * static_cast<double>(effective_steps)));
                tc_index = std::max(0, std::min(tc_index, effective_steps));
                tc_end = std::max(tc_index, std::min(tc_end, effective_steps));
                int taper_len = tc_end - tc_index;
                const double pi_const = 3.14159265358979323846;
                for (int s = 0; s < effective_steps; ++s) {
                    double val = k_smooth[s];
                    if (s < tc_index) {
                    } else if (s < tc_end && taper_len > 0) {
                        double t = (static_cast<double>(s - tc_index)) / static_cast<double>(taper_len);
                        double w = o.taper_final_amplitude +
                                   (1.0 - o.taper_final_amplitude) * 0.5 * (1.0 + std::cos(pi_const * t));
                        val *= w;
                    } else {
                        val = 0.0;
                    }
                    T.Kproc[base + s] = val;
                }
                for (int s = effective_steps; s < steps; ++s) T.Kproc[base + s] = 0.0;
            }
        }
    }
    T.tapered = true;
}

// ------------------------------------------------------------------------------------
// Waves
// ------------------------------------------------------------------------------------
// src/wave_types.cpp:178-255 ComputeWaveNumber
static double ComputeWaveNumber(double omega, double water_depth, double g, double tolerance = 1e-6,
                                int max_iterations = 100) {
    constexpr double DEEP_WATER_THRESHOLD = 1000.0;
    if (omega <= 0.0) throw std::runtime_error("Angular frequency must be positive.");
    if (water_depth < 0.0) throw std::runtime_error("Water depth cannot be negative.");
    if (g <= 0.0) throw std::runtime_error("Gravity must be positive.");
    if (tolerance <= 0.0) throw std::runtime_error("Tolerance must be positive.");
    if (max_iterations <= 0) throw std::runtime_error("Maximum iterations must be positive.");
    if (water_depth == 0.0 || water_depth > DEEP_WATER_THRESHOLD || std::isinf(water_depth)) {
        return omega * omega / g;
    }
    double k = omega * omega / g;
    int iterations = 0;
    double error = 1.0;
    while (error > tolerance && iterations < max_iterations) {
        double tanh_kh = std::tanh(k * water_depth);
        double f = omega * omega - g * k * tanh_kh;
        double df = -2.0 * g * tanh_kh - g * k * water_depth * (1.0 - tanh_kh * tanh_kh);
        if (std::abs(df) < tolerance) throw std::runtime_error("Numerical instability: derivative too close to zero.");
        double delta_k = f / df;
        k -= delta_k;
        error = std::abs(delta_k);
        iterations++;
    }
    if (iterations >= max_iterations) throw std::runtime_error("Failed to converge within maximum iterations.");
    return k;
}

// src/wave_types.cpp:679-693
static Vec PiersonMoskowitzSpectrumHz(Vec& f, double Hs, double Tp) {
    std::sort(f.begin(), f.end());
    Vec S(f.size());
    for (size_t i = 0; i < f.size(); ++i) {
        S[i] = 1.25 * std::pow(1 / Tp, 4) * std::pow(Hs / 2, 2) * std::pow(f[i], -5) *
               std::exp(-1.25 * std::pow(1 / Tp, 4) * std::pow(f[i], -4));
    }
    return S;
}

// src/wave_types.cpp:695-715
static Vec JONSWAPSpectrumHz(Vec& f, double Hs, double Tp, double gamma, bool is_normalized) {
    Vec S = PiersonMoskowitzSpectrumHz(f, Hs, Tp);
    double normalization_factor = (1 - 0.287 * std::log(gamma));
    for (size_t i = 0; i < S.size(); ++i) {
        double sigma = (f[i] <= 1.0 / Tp) ? 0.07 : 0.09;
        S[i] *= std::pow(gamma, std::exp(-(1.0 / (2.0 * std::pow(sigma, 2))) * std::pow(f[i] * Tp - 1.0, 2)));
        if (is_normalized) S[i] *= normalization_factor;
    }
    return S;
}

// std::uniform_real_distribution<double>(a,b)(std::mt19937&) as libstdc++/MSVC implement it:
// generate_canonical<double,53> draws two 32-bit words, (x0 + x1*2^32)/2^64, clamps below 1.
// [SURVEY.md A.4: this convention reproduces the reference's irregular-wave golden.]
static double UniformReal(std::mt19937& rng, double a, double b) {
    const double r = 4294967296.0;  // 2^32
    double sum = double(rng());
    sum += double(rng()) * r;
    double ret = sum / (r * r);
    if (ret >= 1.0) ret = std::nextafter(1.0, 0.0);
    return ret * (b - a) + a;
}

struct IrregularParams {  // include/hydroc/wave_types.h:277-292
    double simulation_dt = 0, simulation_duration = 0, ramp_duration = 0;
    double wave_height = 0, wave_period = 0;
    double frequency_min = 0.001, frequency_max = 1.0;
    double nfrequencies = 0;
    double peak_enhancement_factor = 1.0;
    int is_normalized = 0;
    int seed = 1;
};

// --- Eigen unsupported/Splines restatement (third-party arithmetic, Eigen 3.4.0) --------------
// KnotAveraging + SplineFitting::Interpolate(pts, 3, params) + Spline::operator().
namespace spline {
static int Span(double u, int degree, const Vec& knots) {
    if (u <= knots[0]) return degree;
    const double* b = knots.data() + degree - 1;
    const double* e = knots.data() + knots.size() - degree - 1;
    const double* pos = std::upper_bound(b, e, u);
    return int(pos - knots.data()) - 1;
}
// Piegl & Tiller A2.2
static void Basis(double u, int p, const Vec& U, int i, double* N) {
    double left[8], right[8];
    left[0] = right[0] = 0.0;
    for (int j = 1; j <= p; ++j) { left[j] = u - U[i + 1 - j]; right[j] = U[i + j] - u; }
    N[0] = 1.0;
    for (int j = 1; j <= p; ++j) {
        double saved = 0.0;
        for (int r = 0; r < j; r++) {
            const double tmp = N[r] / (right[r + 1] + left[j - r]);
            N[r] = saved + right[r + 1] * tmp;
            saved = left[j - r] * tmp;
        }
        N[j] = saved;
    }
}
// Interpolate `dim` rows of `pts` ([dim][n], row-major) at parameters `params` ([n]); evaluate at `u_new`.
// Eigen solves the dense collocation system with HouseholderQR; here: dense LU with partial pivoting
// (same solution to rounding; the system is well conditioned).
static void InterpolateAndEval(const double* pts, int dim, int n, const Vec& params, const Vec& u_new, double* out) {
    const int p = 3;
    Vec knots(n + p + 1);
    for (int j = 1; j < n - p; ++j) {
        double s = 0;
        for (int q = 0; q < p; ++q) s += params[j + q];
        knots[j + p] = s / double(p);
    }
    for (int j = 0; j <= p; ++j) { knots[j] = 0.0; knots[knots.size() - 1 - j] = 1.0; }
    std::vector<double> A(size_t(n) * n, 0.0);
    for (int i = 1; i < n - 1; ++i) {
        const int span = Span(params[i], p, knots);
        double Nb[4];
        Basis(params[i], p, knots, span, Nb);
        for (int q = 0; q <= p; ++q) A[size_t(i) * n + span - p + q] = Nb[q];
    }
    A[0] = 1.0;
    A[size_t(n - 1) * n + n - 1] = 1.0;
    // RHS: [n][dim]
    std::vector<double> X(size_t(n) * dim);
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d) X[size_t(i) * dim + d] = pts[size_t(d) * n + i];
    // LU with partial pivoting, exploiting the band (nonzeros within +-4 of the diagonal) for speed only.
    const int bw = 4;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = std::abs(A[size_t(k) * n + k]);
        const int rmax = std::min(n - 1, k + bw);
        for (int r = k + 1; r <= rmax; ++r) {
            double v = std::abs(A[size_t(r) * n + k]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0.0) throw std::runtime_error("spline collocation matrix singular");
        const int cmax = std::min(n - 1, k + 2 * bw);
        if (piv != k) {
            for (int c = k; c <= cmax; ++c) std::swap(A[size_t(k) * n + c], A[size_t(piv) * n + c]);
            for (int d = 0; d < dim; ++d) std::swap(X[size_t(k) * dim + d], X[size_t(piv) * dim + d]);
        }
        for (int r = k + 1; r <= rmax; ++r) {
            const double m = A[size_t(r) * n + k] / A[size_t(k) * n + k];
            if (m == 0.0) continue;
            for (int c = k; c <= cmax; ++c) A[size_t(r) * n + c] -= m * A[size_t(k) * n + c];
            for (int d = 0; d < dim; ++d) X[size_t(r) * dim + d] -= m * X[size_t(k) * dim + d];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        const int cmax = std::min(n - 1, k + 2 * bw);
        for (int d = 0; d < dim; ++d) {
            double s = X[size_t(k) * dim + d];
            for (int c = k + 1; c <= cmax; ++c) s -= A[size_t(k) * n + c] * X[size_t(c) * dim + d];
            X[size_t(k) * dim + d] = s / A[size_t(k) * n + k];
        }
    }
    // evaluate: out [dim][m]
    const int m = (int)u_new.size();
    for (int j = 0; j < m; ++j) {
        const double u = u_new[j];
        const int span = Span(u, p, knots);
        double Nb[4];
        Basis(u, p, knots, span, Nb);
        for (int d = 0; d < dim; ++d) {
            double s = 0.0;
            for (int q = 0; q <= p; ++q) s += Nb[q] * X[size_t(span - p + q) * dim + d];
            out[size_t(d) * m + j] = s;
        }
    }
}
}  // namespace spline

// src/helper.cpp:8-22
static size_t get_lower_index(double value, const Vec& ticks) {
    auto it = std::upper_bound(ticks.begin(), ticks.end(), value);
    size_t idx = it - ticks.begin() - 1;
    if (ticks[idx] == value) idx -= 1;
    if (idx <= 0 || idx >= ticks.size() - 1) {
        throw std::runtime_error("Could not find index for value " + std::to_string(value) + " in array with bounds (" +
                                 std::to_string(ticks.front()) + ", " + std::to_string(ticks.back()) + ").");
    }
    return idx;
}

enum WaveMode { kNoWave = 0, kRegular = 1, kIrregular = 2 };

// Resampled excitation IRF (shared between realisations of one design): wave_types.cpp:432-449,572-628
struct ExcIRF {
    int N = 0;
    std::vector<Vec> t;   // [N][Le]
    std::vector<Vec> w;   // [N][Le]
    std::vector<Vec> f;   // [N][6*Le]  (dof-major)
    std::vector<int> Le;
};

static std::shared_ptr<ExcIRF> BuildExcIRF(const Tables& T, double dt) {
    auto E = std::make_shared<ExcIRF>();
    E->N = T.N;
    E->t.resize(T.N); E->w.resize(T.N); E->f.resize(T.N); E->Le.resize(T.N);
    for (int b = 0; b < T.N; ++b) {
        Vec t_old(T.exc_irf_t.begin() + size_t(b) * T.Le0, T.exc_irf_t.begin() + size_t(b + 1) * T.Le0);
        const double* f_old = T.exc_irf_f.data() + size_t(b) * 6 * T.Le0;
        if (dt > 0.0) {  // wave_types.cpp:447-449, ResampleIRF :572-606
            double t0 = t_old[0], t1 = t_old[t_old.size() - 1];
            Vec t_new = LinSpaced(static_cast<int>(std::ceil((t1 - t0) / dt)), t0, t1);
            Vec u_old = LinSpaced((int)t_old.size(), 0, 1);
            Vec u_new = LinSpaced((int)t_new.size(), 0, 1);
            Vec f_new(size_t(6) * t_new.size());
            spline::InterpolateAndEval(f_old, 6, (int)t_old.size(), u_old, u_new, f_new.data());
            E->t[b] = t_new;
            E->f[b] = f_new;
        } else {
            E->t[b] = t_old;
            E->f[b].assign(f_old, f_old + size_t(6) * T.Le0);
        }
        E->w[b] = WidthArray(E->t[b]);
        E->Le[b] = (int)E->t[b].size();
    }
    return E;
}

struct Waves {
    int mode = kNoWave;
    // regular (wave_types.cpp:266-352)
    double reg_amplitude = 0, reg_omega = 0, reg_phase = 0, wavenumber = 0;
    Vec reg_mag, reg_phase_interp;  // [D]
    // irregular
    IrregularParams ip;
    std::shared_ptr<ExcIRF> irf;
    Vec freqs, S, widths, phases, wavenumbers;
    Vec eta_t, eta;
};

// wave_types.cpp:278-299,329-352 RegularWave::AddH5Data + GetOmegaDelta + Get*Interp; :274-276 Initialize
static void SetupRegular(const Tables& T, Waves& W) {
    const int total_dofs = 6 * T.N;
    W.reg_mag.assign(total_dofs, 0.0);
    W.reg_phase_interp.assign(total_dofs, 0.0);
    double omega_max = T.w_list[T.w_list.size() - 1];
    double num_freqs = (double)T.w_list.size();
    double wave_omega_delta = omega_max / num_freqs;
    double freq_index_des = (W.reg_omega / wave_omega_delta) - 1;
    for (int b = 0; b < T.N; b++) {
        for (int rowEx = 0; rowEx < 6; rowEx++) {
            const double* mag = T.exc_mag.data() + (size_t(b) * 6 + rowEx) * T.nw;
            const double* ph = T.exc_phase.data() + (size_t(b) * 6 + rowEx) * T.nw;
            double fi = freq_index_des - std::floor(freq_index_des);
            int i0 = (int)std::floor(freq_index_des);
            if (i0 < 0 || i0 + 1 >= T.nw) throw std::out_of_range("regular wave omega outside excitation frequency table");
            W.reg_mag[6 * b + rowEx] = (fi * (mag[i0 + 1] - mag[i0])) + mag[i0];
            W.reg_phase_interp[6 * b + rowEx] = (fi * (ph[i0 + 1] - ph[i0])) + ph[i0];
        }
    }
    W.wavenumber = ComputeWaveNumber(W.reg_omega, T.depth, T.g);
}

// wave_types.cpp:643-676 CreateSpectrum
static void CreateSpectrum(const Tables& T, Waves& W) {
    const IrregularParams& p = W.ip;
    int nf;
    if (p.nfrequencies == 0) {
        double df = 1.0 / p.simulation_duration;
        nf = std::ceil((p.frequency_max - p.frequency_min) / df);
    } else {
        nf = p.nfrequencies;
    }
    W.freqs = LinSpaced(nf, p.frequency_min, p.frequency_max);
    W.S = JONSWAPSpectrumHz(W.freqs, p.wave_height, p.wave_period, p.peak_enhancement_factor, p.is_normalized != 0);
    W.widths = WidthArray(W.freqs);
    W.phases.resize(nf);
    std::mt19937 rng(p.seed);
    for (int i = 0; i < nf; ++i) W.phases[i] = UniformReal(rng, 0.0, 2 * M_PI);
    W.wavenumbers.resize(nf);
    for (int i = 0; i < nf; ++i) W.wavenumbers[i] = ComputeWaveNumber(2 * M_PI * W.freqs[i], T.depth, T.g);
}

// wave_types.cpp:14-59 GetEta / GetEtaIrregular at position (0,0,0)
static double EtaIrregular(double x_pos, double time, const Waves& W) {
    double eta = 0.0;
    const size_t nf = W.freqs.size();
    for (size_t i = 0; i < nf; ++i) {
        double amplitude = std::sqrt(2 * W.S[i] * W.widths[i]);
        double omega = 2 * M_PI * W.freqs[i];
        eta += amplitude * std::cos(W.wavenumbers[i] * x_pos - omega * time + W.phases[i]);
    }
    return eta;
}

// ---- water kinematics (off the per-step force path; WaveBase surface) -------------------------------------------
struct V3 { double x = 0, y = 0, z = 0; };
// wave_types.cpp:14-25 GetEta
static double GetEta(const V3& position, double time, double omega, double amplitude, double phase, double wavenumber) {
    double x_pos = position.x;
    double eta = amplitude * std::cos(wavenumber * x_pos - omega * time + phase);
    return eta;
}
// wave_types.cpp:61-91 GetWaterVelocity
static V3 GetWaterVelocity(const V3& position, double time, double omega, double amplitude, double phase, double wavenumber,
                           double water_depth, double mwl) {
    double x_pos = position.x;
    double z_pos = position.z - mwl;
    V3 water_velocity;
    if (2 * M_PI / wavenumber > water_depth || wavenumber * water_depth > 500.0) {
        water_velocity.x = omega * amplitude * std::exp(wavenumber * z_pos) * std::cos(wavenumber * x_pos - omega * time + phase);
        water_velocity.z = omega * amplitude * std::exp(wavenumber * z_pos) * std::sin(wavenumber * x_pos - omega * time + phase);
    } else {
        water_velocity.x = omega * amplitude * std::cosh(wavenumber * (z_pos + water_depth)) /
                           std::sinh(wavenumber * water_depth) * std::cos(wavenumber * x_pos - omega * time + phase);
        water_velocity.z = omega * amplitude * std::sinh(wavenumber * (z_pos + water_depth)) /
                           std::sinh(wavenumber * water_depth) * std::sin(wavenumber * x_pos - omega * time + phase);
    }
    return water_velocity;
}
// wave_types.cpp:93-122 GetWaterAcceleration
static V3 GetWaterAcceleration(const V3& position, double time, double omega, double amplitude, double phase,
                               double wavenumber, double water_depth, double mwl) {
    double x_pos = position.x;
    double z_pos = position.z - mwl;
    V3 water_acceleration;
    if (2 * M_PI / wavenumber > water_depth || wavenumber * water_depth > 500.0) {
        water_acceleration.x =
            omega * omega * amplitude * std::exp(wavenumber * z_pos) * std::sin(wavenumber * x_pos - omega * time + phase);
        water_acceleration.z =
            -omega * omega * amplitude * std::exp(wavenumber * z_pos) * std::cos(wavenumber * x_pos - omega * time + phase);
    } else {
        water_acceleration.x = omega * omega * amplitude * std::cosh(wavenumber * (z_pos + water_depth)) /
                               std::sinh(wavenumber * water_depth) * std::sin(wavenumber * x_pos - omega * time + phase);
        water_acceleration.z = -omega * omega * amplitude * std::sinh(wavenumber * (z_pos + water_depth)) /
                               std::sinh(wavenumber * water_depth) * std::cos(wavenumber * x_pos - omega * time + phase);
    }
    return water_acceleration;
}
// wave_types.cpp:124-141 / 143-160 GetWaterVelocityIrregular / GetWaterAccelerationIrregular
static V3 GetWaterVelocityIrregular(const V3& position, double time, const Waves& W, double water_depth, double mwl) {
    V3 water_velocity;
    for (size_t i = 0; i < W.freqs.size(); ++i) {
        double amplitude = std::sqrt(2 * W.S[i] * W.widths[i]);
        double omega = 2 * M_PI * W.freqs[i];
        V3 c = GetWaterVelocity(position, time, omega, amplitude, W.phases[i], W.wavenumbers[i], water_depth, mwl);
        water_velocity.x += c.x; water_velocity.y += c.y; water_velocity.z += c.z;
    }
    return water_velocity;
}
static V3 GetWaterAccelerationIrregular(const V3& position, double time, const Waves& W, double water_depth, double mwl) {
    V3 water_acceleration;
    for (size_t i = 0; i < W.freqs.size(); ++i) {
        double amplitude = std::sqrt(2 * W.S[i] * W.widths[i]);
        double omega = 2 * M_PI * W.freqs[i];
        V3 c = GetWaterAcceleration(position, time, omega, amplitude, W.phases[i], W.wavenumbers[i], water_depth, mwl);
        water_acceleration.x += c.x; water_acceleration.y += c.y; water_acceleration.z += c.z;
    }
    return water_acceleration;
}
// {NoWave,RegularWave,IrregularWaves}::GetElevation / GetVelocity / GetAcceleration
// (include/hydroc/wave_types.h:103-109, wave_types.cpp:301-313, :515-550 incl. Wheeler stretching)
static void WaveKinematics(const Tables& T, const Waves& W, const V3& position, double time, bool wave_stretching,
                           double mwl, double& eta, V3& vel, V3& acc) {
    eta = 0.0; vel = V3(); acc = V3();
    if (W.mode == kRegular) {
        eta = GetEta(position, time, W.reg_omega, W.reg_amplitude, W.reg_phase, W.wavenumber);
        vel = GetWaterVelocity(position, time, W.reg_omega, W.reg_amplitude, W.reg_phase, W.wavenumber, T.depth, mwl);
        acc = GetWaterAcceleration(position, time, W.reg_omega, W.reg_amplitude, W.reg_phase, W.wavenumber, T.depth, mwl);
    } else if (W.mode == kIrregular) {
        eta = EtaIrregular(position.x, time, W);
        V3 position_stretched = position;
        if (wave_stretching) {
            double z_pos = position.z - mwl;
            position_stretched.z = T.depth * (z_pos - eta) / (T.depth + eta);      // Wheeler stretching
        }
        vel = GetWaterVelocityIrregular(position_stretched, time, W, T.depth, mwl);
        acc = GetWaterAccelerationIrregular(position_stretched, time, W, T.depth, mwl);
    }
}

// wave_types.cpp:717-774 CreateFreeSurfaceElevation
static void CreateFreeSurfaceElevation(Waves& W) {
    const IrregularParams& p = W.ip;
    double t_irf_min = 0.0, t_irf_max = 0.0;
    for (size_t ii = 0; ii < W.irf->t.size(); ii++) {
        const Vec& tt = W.irf->t[ii];
        if (tt[0] < t_irf_min) t_irf_min = tt[0];
        if (tt[0] > t_irf_max) t_irf_max = tt[0];
        if (tt[tt.size() - 1] > t_irf_max) t_irf_max = tt[tt.size() - 1];
        if (tt[tt.size() - 1] < t_irf_min) t_irf_min = tt[tt.size() - 1];
    }
    double duration = p.simulation_duration + 2 * (t_irf_max - t_irf_min);
    int num_timesteps = static_cast<int>(std::ceil(duration / p.simulation_dt));
    W.eta_t = LinSpaced(num_timesteps + 1, 0, num_timesteps * p.simulation_dt);
    for (size_t ii = 0; ii < W.eta_t.size(); ii++) W.eta_t[ii] += -t_irf_max;
    W.eta.resize(W.eta_t.size());
    // (samples are independent; the OpenMP loop only speeds up test set-up, each sample's sum order is unchanged)
#pragma omp parallel for schedule(static)
    for (long j = 0; j < (long)W.eta_t.size(); ++j) W.eta[j] = EtaIrregular(0.0, W.eta_t[j], W);
    if (p.ramp_duration > 0.0) {
        for (size_t i = 0; i < W.eta_t.size(); ++i) {
            if (W.eta_t[i] < p.ramp_duration) {
                if (W.eta_t[i] <= 0.0) W.eta[i] *= 0.0;
                else W.eta[i] *= W.eta_t[i] / p.ramp_duration;
            }
        }
    }
}

// wave_types.cpp:776-844 ExcitationConvolution
static double ExcitationConvolution(const Waves& W, int body, int dof, double time) {
    double f_ex = 0.0;
    const Vec& irf_time_array = W.irf->t[body];
    const double* irf_val = W.irf->f[body].data() + size_t(dof) * W.irf->Le[body];
    const Vec& irf_width_array = W.irf->w[body];
    const Vec& fst = W.eta_t;
    double tmin = fst.front(), tmax = fst.back();
    double t_tau0 = time - irf_time_array[0];
    long idx = 0;
    if (t_tau0 <= tmin) idx = 0;
    else if (t_tau0 >= tmax) idx = (long)fst.size() - 2;
    else idx = (long)get_lower_index(t_tau0, fst);
    for (size_t j = 0; j < irf_time_array.size(); ++j) {
        double tau = irf_time_array[j];
        double t_tau = time - tau;
        if (tmin <= t_tau && t_tau <= tmax) {
            while (fst[idx] > t_tau) idx -= 1;
            double t1 = fst[idx], t2 = fst[idx + 1];
            double eta_val;
            if (t_tau == t1) eta_val = W.eta[idx];
            else if (t_tau == t2) eta_val = W.eta[idx + 1];
            else if (t_tau > t1 && t_tau < t2) {
                double eta1 = W.eta[idx], eta2 = W.eta[idx + 1];
                double w1 = (t2 - t_tau) / (t2 - t1);
                double w2 = 1.0 - w1;
                eta_val = w1 * eta1 + w2 * eta2;
            } else {
                throw std::runtime_error("Excitation convolution: wrong tau value " + std::to_string(tau) +
                                         " not between " + std::to_string(t1) + " and " + std::to_string(t2) + ".");
            }
            f_ex += irf_val[j] * eta_val * irf_width_array[j];
        } else {
            throw std::runtime_error(
                "Excitation convolution: trying to find free surface elevation at a time out of bounds from the "
                "precomputed free surface elevation (" + std::to_string(t_tau) + "not in [" + std::to_string(tmin) +
                ", " + std::to_string(tmax) + "]). Excitation force ignored at this time step.");
        }
    }
    return f_ex;
}

// WaveBase::GetForceAtTime family: wave_types.cpp:257-264 (NoWave), :315-327 (RegularWave), :552-570 (Irregular)
static void WaveForceAtTime(const Tables& T, const Waves& W, double t, double* f) {
    const int D = T.D;
    for (int i = 0; i < D; i++) f[i] = 0.0;
    if (W.mode == kRegular) {
        for (int b = 0; b < T.N; b++) {
            int body_offset = 6 * b;
            for (int rowEx = 0; rowEx < 6; rowEx++) {
                // NB: phase indexed [rowEx] (body 0's phases for every body) -- reference quirk, :323
                f[body_offset + rowEx] = W.reg_mag[body_offset + rowEx] * W.reg_amplitude *
                                         std::cos(W.reg_omega * t + W.reg_phase_interp[rowEx]);
            }
        }
    } else if (W.mode == kIrregular) {
        for (int body = 0; body < T.N; body++)
            for (int dof = 0; dof < 6; ++dof) f[body * 6 + dof] = ExcitationConvolution(W, body, dof, t);
    }
}

// ------------------------------------------------------------------------------------
// TestHydro state for one system instance
// ------------------------------------------------------------------------------------
struct Instance {
    std::shared_ptr<Tables> T;
    Waves W;
    Vec time_history;                         // newest first (hydro_forces.cpp:560)
    std::vector<std::vector<Vec>> vel_hist;   // [N][hist][6], newest first
    Vec f_hs, f_rad, f_wave, f_total;
    double prev_time = -1;                    // hydro_forces.cpp:176
    int omp_mode = 1;                         // 1: OpenMP branch semantics (the build the reference ships), 0: serial #else branch
    double sec_hs = 0, sec_rad = 0, sec_wave = 0;
    std::string err;
};

// hydro_forces.cpp:263-322
static void ComputeForceHydrostatics(Instance& I, const double* pose, const double* gvec) {
    const Tables& T = *I.T;
    const double rho = T.rho;
    const double glen = std::sqrt(gvec[0] * gvec[0] + gvec[1] * gvec[1] + gvec[2] * gvec[2]);  // ChVector3::Length
    const double rho_times_g = rho * glen;
    for (int b = 0; b < T.N; ++b) {
        const int off = kDofPerBody * b;
        double* out = &I.f_hs[off];
        const double* eq = &T.equilibrium[off];
        double disp[6];
        for (int i = 0; i < 6; ++i) disp[i] = pose[off + i] - eq[i];
        const double* Kh = &T.lin[size_t(b) * 36];
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
            for (int j = 0; j < 6; ++j) s += Kh[i * 6 + j] * disp[j];
            out[i] += -rho_times_g * s;
        }
        const double V = T.disp_vol[b];
        double buoy[3];
        for (int i = 0; i < 3; ++i) buoy[i] = (rho * (-gvec[i])) * V;  // rho * (-g) * disp_vol, left to right
        out[0] += buoy[0]; out[1] += buoy[1]; out[2] += buoy[2];
        const double* r = &T.cb_minus_cg[kDofLinOrRot * b];
        out[3] += r[1] * buoy[2] - r[2] * buoy[1];
        out[4] += r[2] * buoy[0] - r[0] * buoy[2];
        out[5] += r[0] * buoy[1] - r[1] * buoy[0];
    }
}

// hydro_forces.cpp:327-340
static void PruneHistory(Instance& I, double history_min_time) {
    Vec& th = I.time_history;
    while (th.size() > 1 && th[th.size() - 2] < history_min_time) {
        th.pop_back();
        for (auto& vb : I.vel_hist) if (!vb.empty()) vb.pop_back();
    }
}
// hydro_forces.cpp:343-371
static void InterpolateVelocity6D(const std::vector<Vec>& vh, size_t newer_index, double q, double older_time,
                                  double newer_time, double out[6]) {
    if (q == older_time) { for (int d = 0; d < 6; ++d) out[d] = vh[newer_index + 1][d]; return; }
    if (q == newer_time) { for (int d = 0; d < 6; ++d) out[d] = vh[newer_index][d]; return; }
    if (q > older_time && q < newer_time) {
        const double time_delta = (newer_time - older_time);
        const double weight_older = (time_delta != 0.0) ? ((newer_time - q) / time_delta) : 0.0;
        const double weight_newer = 1.0 - weight_older;
        for (int d = 0; d < 6; ++d) out[d] = weight_older * vh[newer_index + 1][d] + weight_newer * vh[newer_index][d];
        return;
    }
    throw std::runtime_error("Radiation convolution: interpolation error; query_time not bracketed by history.");
}
// hydro_forces.cpp:374-381
static bool AdvanceToBracket(const Vec& th, size_t& index, double q) {
    while ((index + 1) < th.size() && th[index + 1] > q) ++index;
    return ((index + 1) < th.size());
}

// hydro_forces.cpp:537-691
static void ComputeForceRadiationDampingConv(Instance& I, double simulation_time, const double* vel) {
    const Tables& T = *I.T;
    const int rirf_steps = T.L, total_dofs = T.D;
    const int rirf_last_index = (int)T.rirf_t.size() - 1;
    const double history_min_time = simulation_time - (rirf_last_index >= 0 ? T.rirf_t[rirf_last_index] : 0.0);
    if (!I.time_history.empty() && simulation_time == I.time_history.front())
        throw std::runtime_error("Tried to compute the radiation damping convolution twice within the same time step!");
    I.time_history.insert(I.time_history.begin(), simulation_time);
    for (int b = 0; b < T.N; ++b) {
        Vec v(vel + 6 * b, vel + 6 * b + 6);
        I.vel_hist[b].insert(I.vel_hist[b].begin(), std::move(v));
    }
    PruneHistory(I, history_min_time);
    if (I.time_history.size() <= 1) return;

    const Vec& th = I.time_history;
    double* out = I.f_rad.data();
    if (I.omp_mode) {
#ifdef _OPENMP
        const int num_threads = omp_get_max_threads();
#else
        const int num_threads = 1;
#endif
        std::vector<Vec> thread_locals(num_threads, Vec(total_dofs, 0.0));
#pragma omp parallel
        {
#ifdef _OPENMP
            const int tid = omp_get_thread_num();
#else
            const int tid = 0;
#endif
            Vec& local_out = thread_locals[tid];
            size_t history_index_local = 0;
#pragma omp for schedule(static)
            for (int step = 0; step < rirf_steps; ++step) {
                const double q = simulation_time - T.rirf_t[step];
                size_t time_index = history_index_local;
                if (!AdvanceToBracket(th, time_index, q)) continue;
                history_index_local = time_index;
                const double newer_time = th[history_index_local];
                const double older_time = th[history_index_local + 1];
                for (int body = 0; body < T.N; ++body) {
                    const auto& vh = I.vel_hist[body];
                    if (vh.size() <= history_index_local) continue;
                    double v6[6];
                    InterpolateVelocity6D(vh, history_index_local, q, older_time, newer_time, v6);
                    const double step_width = T.rirf_w[step];
                    if (step_width == 0.0) continue;
                    for (int dof = 0; dof < 6; ++dof) {
                        const int col = body * 6 + dof;
                        const double scale = v6[dof] * step_width;
                        if (scale == 0.0) continue;
                        for (int row = 0; row < total_dofs; ++row) local_out[row] += T.RIRFval(row, col, step) * scale;
                    }
                }
            }
        }
        for (int t = 0; t < num_threads; ++t)
            for (int row = 0; row < total_dofs; ++row) out[row] += thread_locals[t][row];
    } else {
        size_t history_index = 0;
        for (int step = 0; step < rirf_steps; ++step) {
            const double q = simulation_time - T.rirf_t[step];
            size_t time_index = history_index;
            if (!AdvanceToBracket(th, time_index, q)) break;
            history_index = time_index;
            const double newer_time = th[history_index];
            const double older_time = th[history_index + 1];
            for (int body = 0; body < T.N; ++body) {
                const auto& vh = I.vel_hist[body];
                if (vh.size() <= history_index) continue;
                double v6[6];
                InterpolateVelocity6D(vh, history_index, q, older_time, newer_time, v6);
                const double step_width = T.rirf_w[step];
                for (int dof = 0; dof < 6; ++dof) {
                    const int col = body * 6 + dof;
                    const double scale = v6[dof] * step_width;
                    if (scale == 0.0) continue;
                    for (int row = 0; row < total_dofs; ++row) out[row] += T.RIRFval(row, col, step) * scale;
                }
            }
        }
    }
}

static inline double now_s() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
}

// hydro_forces.cpp:727-767 CoordinateFuncForBody (the recompute branch) + :713-725 ComputeForceWaves.
// Returns true when a recompute happened, false when the cached total was reused.
static bool EvaluateAtTime(Instance& I, double t, const double* pose, const double* vel, const double* gvec) {
    if (t == I.prev_time) return false;
    I.prev_time = t;
    const int D = I.T->D;
    std::fill(I.f_total.begin(), I.f_total.end(), 0.0);
    std::fill(I.f_hs.begin(), I.f_hs.end(), 0.0);
    std::fill(I.f_rad.begin(), I.f_rad.end(), 0.0);
    std::fill(I.f_wave.begin(), I.f_wave.end(), 0.0);
    double t0 = now_s();
    ComputeForceHydrostatics(I, pose, gvec);
    double t1 = now_s();
    ComputeForceRadiationDampingConv(I, t, vel);
    double t2 = now_s();
    WaveForceAtTime(*I.T, I.W, t, I.f_wave.data());
    double t3 = now_s();
    I.sec_hs += t1 - t0; I.sec_rad += t2 - t1; I.sec_wave += t3 - t2;
    for (int i = 0; i < D; i++) I.f_total[i] = I.f_hs[i] - I.f_rad[i] + I.f_wave[i];
    return true;
}

}  // namespace orc

// =====================================================================================
// C interface (ctypes)
// =====================================================================================
using namespace orc;

struct OrcTables { std::shared_ptr<Tables> p; };
struct OrcInstance { Instance I; };

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH } catch (const std::out_of_range& e) { g_err = e.what(); return -2; } \
                   catch (const std::exception& e) { g_err = e.what(); return -1; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// Arrays are RAW file values (row-major C order, as in the BEMIO .h5):
//  K [N][6][D][L], lin [N][6][6], ainf [N][6][D], cg/cb [N][3], exc_mag/phase [N][6][nw], exc_irf_f [N][6][Le0],
//  rirf_t [N][L], exc_irf_t [N][Le0].
OrcTables* orc_tables_create(int N, int L, const double* rirf_t, const double* K, double rho, double g, double depth,
                             const double* lin, const double* ainf, const double* disp_vol, const double* cg,
                             const double* cb, int nw, const double* w_list, const double* exc_mag,
                             const double* exc_phase, int Le0, const double* exc_irf_t, const double* exc_irf_f) {
    try {
        auto T = std::make_shared<Tables>();
        T->N = N; T->D = 6 * N; T->L = L; T->nw = nw; T->Le0 = Le0;
        T->rho = rho; T->g = g; T->depth = depth;
        const int D = T->D;
        // HydroData::GetRIRFTimeVector (h5fileinfo.cpp:325-343)
        T->rirf_t.assign(rirf_t, rirf_t + L);
        for (int b = 1; b < N; ++b)
            for (int j = 0; j < L; ++j)
                if (std::abs(rirf_t[size_t(b) * L + j] - rirf_t[j]) > 1e-10)
                    throw std::runtime_error("RIRF time vectors have to be exactly the same for all bodies.");
        T->rirf_w = WidthArray(T->rirf_t);
        T->K.assign(K, K + size_t(N) * 6 * D * L);
        T->lin.assign(lin, lin + size_t(N) * 36);
        T->ainf.assign(ainf, ainf + size_t(N) * 6 * D);
        for (auto& v : T->ainf) v *= rho;
        T->disp_vol.assign(disp_vol, disp_vol + N);
        T->cg.assign(cg, cg + 3 * N);
        T->cb.assign(cb, cb + 3 * N);
        T->equilibrium.assign(D, 0.0);
        T->cb_minus_cg.assign(3 * N, 0.0);
        for (int b = 0; b < N; ++b)
            for (int i = 0; i < 3; ++i) {
                T->equilibrium[i + 6 * b] = T->cg[3 * b + i];
                T->cb_minus_cg[i + 3 * b] = T->cb[3 * b + i] - T->cg[3 * b + i];
            }
        if (nw > 0) {
            T->w_list.assign(w_list, w_list + nw);
            T->exc_mag.assign(exc_mag, exc_mag + size_t(N) * 6 * nw);
            const double rg = rho * g;
            for (auto& v : T->exc_mag) v = v * rg;
            T->exc_phase.assign(exc_phase, exc_phase + size_t(N) * 6 * nw);
        }
        if (Le0 > 0) {
            T->exc_irf_t.assign(exc_irf_t, exc_irf_t + size_t(N) * Le0);
            T->exc_irf_f.assign(exc_irf_f, exc_irf_f + size_t(N) * 6 * Le0);
            const double rg = rho * g;
            for (auto& v : T->exc_irf_f) v *= rg;
        }
        return new OrcTables{T};
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void orc_tables_destroy(OrcTables* t) { delete t; }

int orc_tables_set_tapered(OrcTables* t, int moving_average, int window_length, double rirf_end_time,
                           double start_percent, double end_percent, double final_amplitude) {
    ORC_TRY
    TaperedOpts o; o.smoothing_moving_average = moving_average; o.window_length = window_length;
    o.rirf_end_time = rirf_end_time; o.taper_start_percent = start_percent; o.taper_end_percent = end_percent;
    o.taper_final_amplitude = final_amplitude;
    ProcessRIRF(*t->p, o);
    return 0;
    ORC_CATCH
}
// effective K(row, col, s) incl. rho / tapering: [D][D][L]
void orc_tables_get_rirf(OrcTables* t, double* out) {
    const Tables& T = *t->p;
    for (int r = 0; r < T.D; ++r) for (int c = 0; c < T.D; ++c) for (int s = 0; s < T.L; ++s)
        out[(size_t(r) * T.D + c) * T.L + s] = T.RIRFval(r, c, s);
}
void orc_tables_get_rirf_width(OrcTables* t, double* out) { std::copy(t->p->rirf_w.begin(), t->p->rirf_w.end(), out); }

// src/chloadaddedmass.cpp:12-52: stacked 6N x 6N, zero-padded to n_sys with the block at (0,0).
int orc_added_mass(OrcTables* t, int n_sys, double* M) {
    const Tables& T = *t->p;
    if (n_sys < T.D) { g_err = "n_sys smaller than 6N"; return -1; }
    std::fill(M, M + size_t(n_sys) * n_sys, 0.0);
    for (int i = 0; i < T.N; i++)
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < T.D; ++c) M[size_t(i * 6 + r) * n_sys + c] = T.ainf[(size_t(i) * 6 + r) * T.D + c];
    return 0;
}
// src/chloadaddedmass.cpp:55-71: R += c * M * w
int orc_added_mass_mv(OrcTables* t, int n_sys, double c, const double* w, double* R) {
    std::vector<double> M(size_t(n_sys) * n_sys);
    if (orc_added_mass(t, n_sys, M.data())) return -1;
    for (int i = 0; i < n_sys; ++i) {
        double s = 0.0;
        for (int j = 0; j < n_sys; ++j) s += (c * M[size_t(i) * n_sys + j]) * w[j];
        R[i] += s;
    }
    return 0;
}

OrcInstance* orc_instance_create(OrcTables* t, int omp_mode) {
    auto* o = new OrcInstance();
    Instance& I = o->I;
    I.T = t->p;
    I.omp_mode = omp_mode;
    const int D = I.T->D;
    I.vel_hist.assign(I.T->N, {});
    I.f_hs.assign(D, 0.0); I.f_rad.assign(D, 0.0); I.f_wave.assign(D, 0.0); I.f_total.assign(D, 0.0);
    I.W.mode = kNoWave;
    return o;
}
void orc_instance_destroy(OrcInstance* o) { delete o; }

int orc_set_nowave(OrcInstance* o) { o->I.W = Waves(); o->I.W.mode = kNoWave; return 0; }

int orc_set_regular(OrcInstance* o, double amplitude, double omega, double phase) {
    ORC_TRY
    Waves W; W.mode = kRegular; W.reg_amplitude = amplitude; W.reg_omega = omega; W.reg_phase = phase;
    SetupRegular(*o->I.T, W);
    o->I.W = std::move(W);
    return 0;
    ORC_CATCH
}

// share_irf_from: another instance whose resampled IRF may be reused (same tables, same dt), or NULL.
int orc_set_irregular(OrcInstance* o, double dt, double duration, double ramp, double Hs, double Tp, double fmin,
                      double fmax, double nfreq, double gamma, int is_normalized, int seed,
                      OrcInstance* share_irf_from) {
    ORC_TRY
    Waves W; W.mode = kIrregular;
    W.ip.simulation_dt = dt; W.ip.simulation_duration = duration; W.ip.ramp_duration = ramp;
    W.ip.wave_height = Hs; W.ip.wave_period = Tp; W.ip.frequency_min = fmin; W.ip.frequency_max = fmax;
    W.ip.nfrequencies = nfreq; W.ip.peak_enhancement_factor = gamma; W.ip.is_normalized = is_normalized;
    W.ip.seed = seed;
    if (share_irf_from && share_irf_from->I.W.irf && share_irf_from->I.W.ip.simulation_dt == dt &&
        share_irf_from->I.T == o->I.T)
        W.irf = share_irf_from->I.W.irf;
    else
        W.irf = BuildExcIRF(*o->I.T, dt);
    if (Hs != 0.0 && Tp != 0.0) {  // wave_types.cpp:454-458
        CreateSpectrum(*o->I.T, W);
        CreateFreeSurfaceElevation(W);
    }
    o->I.W = std::move(W);
    return 0;
    ORC_CATCH
}

// Free-surface elevation imported as a (time, eta) series instead of synthesised from a spectrum
// (IrregularWaves::ReadEtaFromFile, wave_types.cpp:480-500, called from InitializeIRFVectors :451-453).  The reference
// snapshot never fills free_surface_time_sampled_ on this branch although ExcitationConvolution reads it (:784-785);
// the intended grid is the file's time column (time_data_), which is what is used here.
int orc_set_irregular_series(OrcInstance* o, double dt, int n, const double* time, const double* eta,
                             OrcInstance* share_irf_from) {
    ORC_TRY
    if (n < 2) throw std::runtime_error("eta series needs at least two samples");
    Waves W; W.mode = kIrregular;
    W.ip.simulation_dt = dt;
    if (share_irf_from && share_irf_from->I.W.irf && share_irf_from->I.W.ip.simulation_dt == dt &&
        share_irf_from->I.T == o->I.T)
        W.irf = share_irf_from->I.W.irf;
    else
        W.irf = BuildExcIRF(*o->I.T, dt);
    W.eta_t.assign(time, time + n);
    W.eta.assign(eta, eta + n);
    o->I.W = std::move(W);
    return 0;
    ORC_CATCH
}

int orc_irregular_sizes(OrcInstance* o, int* nf, int* n_eta, int* Le /*[N]*/) {
    const Waves& W = o->I.W;
    if (W.mode != kIrregular) return -1;
    *nf = (int)W.freqs.size(); *n_eta = (int)W.eta.size();
    for (int b = 0; b < o->I.T->N; ++b) Le[b] = W.irf->Le[b];
    return 0;
}
// any pointer may be NULL
int orc_irregular_get(OrcInstance* o, double* freqs, double* S, double* widths, double* phases, double* wavenumbers,
                      double* eta_t, double* eta) {
    const Waves& W = o->I.W;
    if (W.mode != kIrregular) return -1;
    auto cp = [](const Vec& v, double* d) { if (d) std::copy(v.begin(), v.end(), d); };
    cp(W.freqs, freqs); cp(W.S, S); cp(W.widths, widths); cp(W.phases, phases); cp(W.wavenumbers, wavenumbers);
    cp(W.eta_t, eta_t); cp(W.eta, eta);
    return 0;
}
int orc_irregular_get_irf(OrcInstance* o, int body, double* t, double* w, double* f /*[6][Le]*/) {
    const Waves& W = o->I.W;
    if (W.mode != kIrregular) return -1;
    auto cp = [](const Vec& v, double* d) { if (d) std::copy(v.begin(), v.end(), d); };
    cp(W.irf->t[body], t); cp(W.irf->w[body], w); cp(W.irf->f[body], f);
    return 0;
}
int orc_regular_get(OrcInstance* o, double* mag, double* phase, double* wavenumber) {
    const Waves& W = o->I.W;
    if (W.mode != kRegular) return -1;
    std::copy(W.reg_mag.begin(), W.reg_mag.end(), mag);
    std::copy(W.reg_phase_interp.begin(), W.reg_phase_interp.end(), phase);
    *wavenumber = W.wavenumber;
    return 0;
}

// One force evaluation, TestHydro::CoordinateFuncForBody semantics (hydro_forces.cpp:727-767):
// recompute on a new time value, otherwise return the cached totals.  pose = (x,y,z,roll,pitch,yaw) per body
// (Cardan XYZ angles; the quaternion conversion is Chrono's and stays on the caller's side), vel = (GetPosDt,
// GetAngVelParent) per body, gvec = system gravity vector.  Outputs may be NULL.
// returns 0 = recomputed, 1 = cache hit, <0 = error (orc_last_error()).
int orc_force(OrcInstance* o, double t, const double* pose, const double* vel, const double* gvec, double* total,
              double* hs, double* rad, double* wave) {
    ORC_TRY
    Instance& I = o->I;
    bool re = EvaluateAtTime(I, t, pose, vel, gvec);
    const int D = I.T->D;
    if (total) std::copy(I.f_total.begin(), I.f_total.begin() + D, total);
    if (hs) std::copy(I.f_hs.begin(), I.f_hs.begin() + D, hs);
    if (rad) std::copy(I.f_rad.begin(), I.f_rad.begin() + D, rad);
    if (wave) std::copy(I.f_wave.begin(), I.f_wave.begin() + D, wave);
    return re ? 0 : 1;
    ORC_CATCH
}

int orc_history_len(OrcInstance* o) { return (int)o->I.time_history.size(); }
void orc_profile(OrcInstance* o, double* sec3) { sec3[0] = o->I.sec_hs; sec3[1] = o->I.sec_rad; sec3[2] = o->I.sec_wave; }

// Standalone pieces, exposed for unit parity tests -------------------------------------------------
double orc_wave_number(double omega, double depth, double g) {
    try { return ComputeWaveNumber(omega, depth, g); } catch (const std::exception& e) { g_err = e.what(); return std::nan(""); }
}
void orc_jonswap(int n, const double* f, double Hs, double Tp, double gamma, int is_normalized, double* S) {
    Vec ff(f, f + n);
    Vec s = JONSWAPSpectrumHz(ff, Hs, Tp, gamma, is_normalized != 0);
    std::copy(s.begin(), s.end(), S);
}
void orc_pierson_moskowitz(int n, const double* f, double Hs, double Tp, double* S) {
    Vec ff(f, f + n);
    Vec s = PiersonMoskowitzSpectrumHz(ff, Hs, Tp);
    std::copy(s.begin(), s.end(), S);
}
void orc_linspaced(int n, double lo, double hi, double* out) { Vec v = LinSpaced(n, lo, hi); std::copy(v.begin(), v.end(), out); }
void orc_phases(int seed, int n, double* out) {
    std::mt19937 rng(seed);
    for (int i = 0; i < n; ++i) out[i] = UniformReal(rng, 0.0, 2 * M_PI);
}
// std::uniform_real_distribution itself, to check UniformReal() against the host C++ library.
void orc_phases_stdlib(int seed, int n, double* out) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> dist(0.0, 2 * M_PI);
    for (int i = 0; i < n; ++i) out[i] = dist(rng);
}
int orc_spline_resample(int dim, int n_old, const double* pts, int n_new, double* out) {
    ORC_TRY
    Vec u_old = LinSpaced(n_old, 0, 1), u_new = LinSpaced(n_new, 0, 1);
    spline::InterpolateAndEval(pts, dim, n_old, u_old, u_new, out);
    return 0;
    ORC_CATCH
}
long orc_get_lower_index(double value, int n, const double* ticks) {
    try { Vec t(ticks, ticks + n); return (long)get_lower_index(value, t); }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// --------------------------------------------------------------------------------------------------
// CPU baseline driver (bench.py cpu_baseline / --impl reference): steps `count` instances through
// `nsteps` force evaluations with a prescribed synthetic motion (state-independent of the force, so
// GPU and CPU arms see identical inputs), returns elapsed seconds.  mode 0 = "reference-style":
// instances one after another, OpenMP across lags inside each radiation call (hydro_forces.cpp:593-647);
// mode 1 = "best-effort": instances distributed one per thread, serial inside.
// pose/vel for instance i at step n: amp[d]*sin(om[d]*t + i*0.01), derivative for vel.
// --------------------------------------------------------------------------------------------------
double orc_bench_steps(OrcInstance** inst, int count, int nsteps, double t0, double dt, int mode,
                       const double* amp, const double* om, const double* gvec, double* checksum) {
    if (count <= 0) return 0.0;
    const int D = inst[0]->I.T->D;
    double cs = 0.0;
    const double tstart = now_s();
    if (mode == 0) {
        std::vector<double> pose(D), vel(D);
        for (int i = 0; i < count; ++i) inst[i]->I.omp_mode = 1;
        for (int n = 0; n < nsteps; ++n) {
            const double t = t0 + n * dt;
            for (int i = 0; i < count; ++i) {
                for (int d = 0; d < D; ++d) {
                    pose[d] = amp[d] * std::sin(om[d] * t + i * 0.01);
                    vel[d] = amp[d] * om[d] * std::cos(om[d] * t + i * 0.01);
                }
                EvaluateAtTime(inst[i]->I, t, pose.data(), vel.data(), gvec);
                cs += inst[i]->I.f_total[2];
            }
        }
    } else {
        for (int i = 0; i < count; ++i) inst[i]->I.omp_mode = 0;
#pragma omp parallel for schedule(static) reduction(+ : cs)
        for (int i = 0; i < count; ++i) {
            std::vector<double> pose(D), vel(D);
            for (int n = 0; n < nsteps; ++n) {
                const double t = t0 + n * dt;
                for (int d = 0; d < D; ++d) {
                    pose[d] = amp[d] * std::sin(om[d] * t + i * 0.01);
                    vel[d] = amp[d] * om[d] * std::cos(om[d] * t + i * 0.01);
                }
                EvaluateAtTime(inst[i]->I, t, pose.data(), vel.data(), gvec);
                cs += inst[i]->I.f_total[2];
            }
        }
    }
    const double el = now_s() - tstart;
    if (checksum) *checksum = cs;
    return el;
}

// WaveBase::GetElevation / GetVelocity / GetAcceleration of the instance's wave object at a point.
int orc_wave_kinematics(OrcInstance* o, const double* position, double time, int wave_stretching, double mwl, double* eta,
                        double* vel, double* acc) {
    ORC_TRY
    V3 p; p.x = position[0]; p.y = position[1]; p.z = position[2];
    double e; V3 v, a;
    WaveKinematics(*o->I.T, o->I.W, p, time, wave_stretching != 0, mwl, e, v, a);
    if (eta) *eta = e;
    if (vel) { vel[0] = v.x; vel[1] = v.y; vel[2] = v.z; }
    if (acc) { acc[0] = a.x; acc[1] = a.y; acc[2] = a.z; }
    return 0;
    ORC_CATCH
}

// Test/bench infrastructure (not a restatement of reference code): loads a velocity history as if the instance had
// been stepped through `n` earlier evaluations (times newest first, vel[n][D]); what hydro_forces.cpp:559-577 would
// have left behind, without paying for n convolutions.  The next orc_force call continues from there.
int orc_set_history(OrcInstance* o, int n, const double* times_newest_first, const double* vel) {
    ORC_TRY
    Instance& I = o->I;
    const int N = I.T->N, D = I.T->D;
    I.time_history.assign(times_newest_first, times_newest_first + n);
    for (int b = 0; b < N; ++b) {
        I.vel_hist[b].resize(n);
        for (int i = 0; i < n; ++i) I.vel_hist[b][i].assign(vel + size_t(i) * D + 6 * b, vel + size_t(i) * D + 6 * b + 6);
    }
    I.prev_time = n > 0 ? times_newest_first[0] : -1.0;
    return 0;
    ORC_CATCH
}

// CPU baseline driver with the GPU arm's exact inputs: step n (time times[n]) takes pose/vel from buffer
// (buf0 + n) % nbuf of pose[nbuf][count][D] / vel[nbuf][count][D].  mode as orc_bench_steps.  When forces != null the
// totals of every step are stored as forces[nsteps][count][D].
double orc_bench_lockstep(OrcInstance** inst, int count, int nsteps, const double* times, int nbuf, int buf0,
                          const double* pose, const double* vel, int mode, const double* gvec, double* forces,
                          double* checksum) {
    if (count <= 0) return 0.0;
    const int D = inst[0]->I.T->D;
    double cs = 0.0;
    const double tstart = now_s();
    auto one = [&](int i, int n) {
        const size_t off = (size_t((buf0 + n) % nbuf) * count + i) * D;
        EvaluateAtTime(inst[i]->I, times[n], pose + off, vel + off, gvec);
        if (forces) std::copy(inst[i]->I.f_total.begin(), inst[i]->I.f_total.end(), forces + (size_t(n) * count + i) * D);
        return inst[i]->I.f_total[2];
    };
    if (mode == 0) {
        for (int i = 0; i < count; ++i) inst[i]->I.omp_mode = 1;
        for (int n = 0; n < nsteps; ++n)
            for (int i = 0; i < count; ++i) cs += one(i, n);
    } else {
        for (int i = 0; i < count; ++i) inst[i]->I.omp_mode = 0;
#pragma omp parallel for schedule(static) reduction(+ : cs)
        for (int i = 0; i < count; ++i)
            for (int n = 0; n < nsteps; ++n) cs += one(i, n);
    }
    const double el = now_s() - tstart;
    if (checksum) *checksum = cs;
    return el;
}

}  // extern "C"
