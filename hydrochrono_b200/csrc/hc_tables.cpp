// Hydro tables: the host half of H5FileInfo/HydroData and of the table set-up in TestHydro's
// constructor.  Reference: src/h5fileinfo.cpp:27-91,309-343; src/hydro_forces.cpp:170-216,385-535;
// src/chloadaddedmass.cpp:12-52.
#include <algorithm>
#include <cstring>

#include "hc_internal.h"

namespace hc {

static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
const char* last_error_cstr() { return g_last_error.c_str(); }

// GetWidthArray (src/wave_types.cpp:608-620) == rirf_width_vector (src/hydro_forces.cpp:181-190)
dvec trapezoid_widths(const dvec& x) {
    const size_t n = x.size();
    dvec w(n, 0.0);
    for (size_t i = 0; i < n; ++i) {
        double acc = 0.0;
        if (i + 1 < n) acc += 0.5 * std::fabs(x[i + 1] - x[i]);
        if (i > 0) acc += 0.5 * std::fabs(x[i] - x[i - 1]);
        w[i] = acc;
    }
    return w;
}

hc_tables* tables_from_desc(const hc_tables_desc& d) {
    if (d.num_bodies < 1) fail(HC_ERR_INVALID, "num_bodies must be >= 1");
    if (d.rirf_steps < 2) fail(HC_ERR_INVALID, "rirf_steps must be >= 2");
    if (!d.rirf_t || !d.rirf_K || !d.lin_matrix || !d.inf_added_mass || !d.disp_vol || !d.cg || !d.cb)
        fail(HC_ERR_INVALID, "missing table pointer");
    auto* T = new hc_tables();
    const int N = d.num_bodies, D = 6 * N, L = d.rirf_steps;
    T->N = N; T->D = D; T->L = L; T->nw = d.num_freqs; T->Le0 = d.exc_irf_steps;
    T->rho = d.rho; T->g = d.g; T->depth = d.water_depth;
    T->body.resize(N);
    const double rho_g = d.rho * d.g;
    for (int b = 0; b < N; ++b) {
        BodyTables& B = T->body[b];
        B.disp_vol = d.disp_vol[b];
        B.rirf_t.assign(d.rirf_t + size_t(b) * L, d.rirf_t + size_t(b + 1) * L);
        B.cg.assign(d.cg + 3 * b, d.cg + 3 * b + 3);
        B.cb.assign(d.cb + 3 * b, d.cb + 3 * b + 3);
        B.lin.assign(d.lin_matrix + 36 * size_t(b), d.lin_matrix + 36 * size_t(b + 1));
        B.ainf.assign(d.inf_added_mass + size_t(b) * 6 * D, d.inf_added_mass + size_t(b + 1) * 6 * D);
        for (double& v : B.ainf) v *= d.rho;                         // h5fileinfo.cpp:60-61
        B.K.assign(d.rirf_K + size_t(b) * 6 * D * L, d.rirf_K + size_t(b + 1) * 6 * D * L);
        if (d.num_freqs > 0) {
            if (!d.w || !d.exc_mag || !d.exc_phase) { delete T; fail(HC_ERR_INVALID, "missing excitation mag/phase"); }
            B.exc_mag.assign(d.exc_mag + size_t(b) * 6 * d.num_freqs, d.exc_mag + size_t(b + 1) * 6 * d.num_freqs);
            for (double& v : B.exc_mag) v = v * rho_g;               // h5fileinfo.cpp:73-75
            B.exc_phase.assign(d.exc_phase + size_t(b) * 6 * d.num_freqs,
                               d.exc_phase + size_t(b + 1) * 6 * d.num_freqs);
        }
        if (d.exc_irf_steps > 0) {
            if (!d.exc_irf_t || !d.exc_irf_f) { delete T; fail(HC_ERR_INVALID, "missing excitation IRF"); }
            B.exc_irf_t.assign(d.exc_irf_t + size_t(b) * d.exc_irf_steps, d.exc_irf_t + size_t(b + 1) * d.exc_irf_steps);
            B.exc_irf_f.assign(d.exc_irf_f + size_t(b) * 6 * d.exc_irf_steps,
                               d.exc_irf_f + size_t(b + 1) * 6 * d.exc_irf_steps);
            for (double& v : B.exc_irf_f) v *= rho_g;                // h5fileinfo.cpp:90
        }
    }
    if (d.num_freqs > 0) T->w_list.assign(d.w, d.w + d.num_freqs);
    // HydroData::GetRIRFTimeVector: all bodies must share the time vector to 1e-10 (h5fileinfo.cpp:325-343)
    T->rirf_t = T->body[0].rirf_t;
    for (int b = 1; b < N; ++b)
        for (int j = 0; j < L; ++j)
            if (std::fabs(T->body[b].rirf_t[j] - T->rirf_t[j]) > 1e-10) {
                delete T;
                fail(HC_ERR_INVALID, "RIRF time vectors have to be exactly the same for all bodies. Difference found in body " +
                                         std::to_string(j) + " at time index " + std::to_string(j) + ".");
            }
    T->rirf_w = trapezoid_widths(T->rirf_t);
    T->equilibrium.assign(D, 0.0);
    T->cb_minus_cg.assign(3 * N, 0.0);
    for (int b = 0; b < N; ++b)
        for (int i = 0; i < 3; ++i) {                                // hydro_forces.cpp:208-216
            T->equilibrium[6 * b + i] = T->body[b].cg[i];
            T->cb_minus_cg[3 * b + i] = T->body[b].cb[i] - T->body[b].cg[i];
        }
    T->rebuild_effective_kernel(nullptr);
    return T;
}

}  // namespace hc

// K_eff(row, col, s): what TestHydro::GetRIRFval returns (hydro_forces.cpp:693-711).  Baseline: raw * rho on
// access (h5fileinfo.cpp:321-323) -- multiplying once here is bitwise identical.  TaperedDirect:
// EnsureProcessedRIRF (hydro_forces.cpp:385-535).
void hc_tables::rebuild_effective_kernel(const hc::TaperOpts* tp) {
    Keff.assign(size_t(D) * D * L, 0.0);
    std::vector<double> raw(L), sm(L);
    for (int row = 0; row < D; ++row) {
        const int b = row / 6, rd = row % 6;
        for (int col = 0; col < D; ++col) {
            const double* src = &body[b].K[(size_t(rd) * D + col) * L];
            double* dst = &Keff[(size_t(row) * D + col) * L];
            if (!tp) {
                for (int s = 0; s < L; ++s) dst[s] = src[s] * rho;
                continue;
            }
            int n_eff = L;
            if (tp->rirf_end_time > 0.0) {
                const double dt = rirf_t[1] - rirf_t[0];
                n_eff = std::min(static_cast<int>(std::floor(tp->rirf_end_time / dt)), L);
            }
            for (int s = 0; s < n_eff; ++s) raw[s] = src[s] * rho;
            if (tp->moving_average) {
                const int half = std::max(3, tp->window_length) / 2;
                for (int s = 0; s < n_eff; ++s) {
                    const int lo = std::max(0, s - half), hi = std::min(n_eff - 1, s + half);
                    double sum = 0.0;
                    for (int i = lo; i <= hi; ++i) sum += raw[i];
                    const int cnt = hi - lo + 1;
                    sm[s] = cnt > 0 ? sum / cnt : raw[s];
                }
            } else if (n_eff >= 5) {
                static const double c5[5] = {-3.0 / 35.0, 12.0 / 35.0, 17.0 / 35.0, 12.0 / 35.0, -3.0 / 35.0};
                sm[0] = raw[0]; sm[1] = raw[1];
                for (int s = 2; s <= n_eff - 3; ++s)
                    sm[s] = c5[0] * raw[s - 2] + c5[1] * raw[s - 1] + c5[2] * raw[s] + c5[3] * raw[s + 1] + c5[4] * raw[s + 2];
                sm[n_eff - 2] = raw[n_eff - 2]; sm[n_eff - 1] = raw[n_eff - 1];
            } else {
                for (int s = 0; s < n_eff; ++s) sm[s] = raw[s];
            }
            int i_start = static_cast<int>(std::floor(tp->start_percent * static_cast<double>(n_eff)));
            int i_end = static_cast<int>(std::floor(tp->end_percent * static_cast<double>(n_eff)));
            i_start = std::max(0, std::min(i_start, n_eff));
            i_end = std::max(i_start, std::min(i_end, n_eff));
            const int span = i_end - i_start;
            const double pi = 3.14159265358979323846;
            for (int s = 0; s < n_eff; ++s) {
                double v = sm[s];
                if (s >= i_start) {
                    if (s < i_end && span > 0) {
                        const double x = static_cast<double>(s - i_start) / static_cast<double>(span);
                        v *= tp->final_amplitude + (1.0 - tp->final_amplitude) * 0.5 * (1.0 + std::cos(pi * x));
                    } else {
                        v = 0.0;
                    }
                }
                dst[s] = v;
            }
            for (int s = n_eff; s < L; ++s) dst[s] = 0.0;
        }
    }
}

// =====================================================================================
// C ABI: tables
// =====================================================================================
namespace hc { const char* last_error_cstr(); }

#define HC_GUARD_BEGIN try {
#define HC_GUARD_END                                                                       \
    }                                                                                      \
    catch (const hc::StatusError& e) { hc::set_last_error(e.msg); return e.code; }         \
    catch (const std::out_of_range& e) { hc::set_last_error(e.what()); return HC_ERR_OUT_OF_RANGE; } \
    catch (const std::exception& e) { hc::set_last_error(e.what()); return HC_ERR_INVALID; }

extern "C" {

const char* hc_last_error(void) { return hc::last_error_cstr(); }
const char* hc_version(void) { return "hydrochrono_b200 0.1.0 (sm_100a)"; }

hc_status hc_tables_create(const hc_tables_desc* desc, hc_tables** out) {
    HC_GUARD_BEGIN
    if (!desc || !out) hc::fail(HC_ERR_INVALID, "null argument");
    *out = hc::tables_from_desc(*desc);
    return HC_OK;
    HC_GUARD_END
}

hc_status hc_tables_load_h5(const char* path, int num_bodies, hc_tables** out) {
    HC_GUARD_BEGIN
    if (!path || !out) hc::fail(HC_ERR_INVALID, "null argument");
    *out = hc::load_bemio_h5(path, num_bodies);
    return HC_OK;
    HC_GUARD_END
}

void hc_tables_destroy(hc_tables* t) { delete t; }

int hc_tables_num_bodies(const hc_tables* t) { return t->N; }
int hc_tables_rirf_steps(const hc_tables* t) { return t->L; }
int hc_tables_num_freqs(const hc_tables* t) { return t->nw; }
int hc_tables_exc_irf_steps(const hc_tables* t) { return t->Le0; }
double hc_tables_rho(const hc_tables* t) { return t->rho; }
double hc_tables_g(const hc_tables* t) { return t->g; }
double hc_tables_water_depth(const hc_tables* t) { return t->depth; }

hc_status hc_tables_rirf_time(const hc_tables* t, double* out) {
    std::copy(t->rirf_t.begin(), t->rirf_t.end(), out);
    return HC_OK;
}
hc_status hc_tables_rirf_width(const hc_tables* t, double* out) {
    std::copy(t->rirf_w.begin(), t->rirf_w.end(), out);
    return HC_OK;
}
hc_status hc_tables_rirf_val(const hc_tables* t, int row, int col, int st, double* out) {
    HC_GUARD_BEGIN
    if (row < 0 || row >= t->D || col < 0 || col >= t->D || st < 0 || st >= t->L)
        throw std::out_of_range("rirfval index out of range in TestHydro");   // hydro_forces.cpp:694-697
    *out = t->Keff[(size_t(row) * t->D + col) * t->L + st];
    return HC_OK;
    HC_GUARD_END
}
hc_status hc_tables_rirf_all(const hc_tables* t, double* out) {
    std::copy(t->Keff.begin(), t->Keff.end(), out);
    return HC_OK;
}
static hc_status check_body(const hc_tables* t, int b) {
    if (b < 0 || b >= t->N) { hc::set_last_error("body index out of range"); return HC_ERR_OUT_OF_RANGE; }
    return HC_OK;
}
hc_status hc_tables_lin_matrix(const hc_tables* t, int b, double* out) {
    if (hc_status s = check_body(t, b)) return s;
    std::copy(t->body[b].lin.begin(), t->body[b].lin.end(), out);
    return HC_OK;
}
hc_status hc_tables_hydrostatic_stiffness(const hc_tables* t, int b, int i, int j, double* out) {
    if (hc_status s = check_body(t, b)) return s;
    if (i < 0 || i > 5 || j < 0 || j > 5) { hc::set_last_error("index out of range"); return HC_ERR_OUT_OF_RANGE; }
    *out = t->body[b].lin[i * 6 + j] * t->rho * t->g;   // h5fileinfo.cpp:313-315
    return HC_OK;
}
hc_status hc_tables_inf_added_mass(const hc_tables* t, int b, double* out) {
    if (hc_status s = check_body(t, b)) return s;
    std::copy(t->body[b].ainf.begin(), t->body[b].ainf.end(), out);
    return HC_OK;
}
hc_status hc_tables_disp_vol(const hc_tables* t, int b, double* out) {
    if (hc_status s = check_body(t, b)) return s;
    *out = t->body[b].disp_vol;
    return HC_OK;
}
hc_status hc_tables_cg(const hc_tables* t, int b, double* out) {
    if (hc_status s = check_body(t, b)) return s;
    std::copy(t->body[b].cg.begin(), t->body[b].cg.end(), out);
    return HC_OK;
}
hc_status hc_tables_cb(const hc_tables* t, int b, double* out) {
    if (hc_status s = check_body(t, b)) return s;
    std::copy(t->body[b].cb.begin(), t->body[b].cb.end(), out);
    return HC_OK;
}

hc_status hc_tables_freq_list(const hc_tables* t, double* out) {
    std::copy(t->w_list.begin(), t->w_list.end(), out);
    return HC_OK;
}
hc_status hc_tables_excitation_mag(const hc_tables* t, int b, double* out) {
    if (hc_status s = check_body(t, b)) return s;
    std::copy(t->body[b].exc_mag.begin(), t->body[b].exc_mag.end(), out);
    return HC_OK;
}
hc_status hc_tables_excitation_phase(const hc_tables* t, int b, double* out) {
    if (hc_status s = check_body(t, b)) return s;
    std::copy(t->body[b].exc_phase.begin(), t->body[b].exc_phase.end(), out);
    return HC_OK;
}
hc_status hc_tables_excitation_irf(const hc_tables* t, int b, double* time, double* f) {
    if (hc_status s = check_body(t, b)) return s;
    if (time) std::copy(t->body[b].exc_irf_t.begin(), t->body[b].exc_irf_t.end(), time);
    if (f) std::copy(t->body[b].exc_irf_f.begin(), t->body[b].exc_irf_f.end(), f);
    return HC_OK;
}

hc_status hc_tables_set_convolution_mode(hc_tables* t, int mode, const hc_tapered_opts* o) {
    HC_GUARD_BEGIN
    if (mode == 0) {
        t->conv_mode = 0;
        t->rebuild_effective_kernel(nullptr);
        return HC_OK;
    }
    if (mode != 1) hc::fail(HC_ERR_INVALID, "unknown convolution mode");
    hc::TaperOpts tp;
    if (o) {
        tp.moving_average = o->smoothing && std::strcmp(o->smoothing, "moving_average") == 0;
        tp.window_length = o->window_length;
        tp.rirf_end_time = o->rirf_end_time;
        tp.start_percent = o->taper_start_percent;
        tp.end_percent = o->taper_end_percent;
        tp.final_amplitude = o->taper_final_amplitude;
    }
    t->conv_mode = 1;
    t->rebuild_effective_kernel(&tp);
    return HC_OK;
    HC_GUARD_END
}

hc_status hc_added_mass(const hc_tables* t, int n_sys, double* M) {
    HC_GUARD_BEGIN
    if (n_sys < t->D) hc::fail(HC_ERR_INVALID, "n_sys must be >= 6 * num_bodies");
    std::fill(M, M + size_t(n_sys) * n_sys, 0.0);
    for (int b = 0; b < t->N; ++b)                                    // chloadaddedmass.cpp:18-24,35-44
        for (int r = 0; r < 6; ++r)
            std::copy(&t->body[b].ainf[size_t(r) * t->D], &t->body[b].ainf[size_t(r + 1) * t->D],
                      &M[size_t(6 * b + r) * n_sys]);
    return HC_OK;
    HC_GUARD_END
}

}  // extern "C"
