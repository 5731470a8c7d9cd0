// WaveBase family (reference src/wave_types.cpp).  Force evaluation is delegated to the device ensemble the wave
// is bound to; wave kinematics (Airy theory, off the per-step path) are evaluated here on the host.
#include <hydroc/wave_types.h>

#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "hc_check.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// ---------------------------------------------------------------------------------------------
// spectra (src/wave_types.cpp:679-715)
// ---------------------------------------------------------------------------------------------
Eigen::VectorXd PiersonMoskowitzSpectrumHz(Eigen::VectorXd& f, double Hs, double Tp) {
    Eigen::VectorXd S(f.size());
    hc_throw_on_error(hc_pierson_moskowitz_spectrum_hz(int(f.size()), f.data(), Hs, Tp, S.data()));
    return S;
}
Eigen::VectorXd JONSWAPSpectrumHz(Eigen::VectorXd& f, double Hs, double Tp, double gamma, bool is_normalized) {
    Eigen::VectorXd S(f.size());
    hc_throw_on_error(hc_jonswap_spectrum_hz(int(f.size()), f.data(), Hs, Tp, gamma, is_normalized ? 1 : 0, S.data()));
    return S;
}

// ---------------------------------------------------------------------------------------------
// Airy kinematics (src/wave_types.cpp:14-160,515-550): the arithmetic lives behind the C ABI (hc_wave_kinematics,
// host code, off the step path as in the reference); waves travel along global +x
// ---------------------------------------------------------------------------------------------
namespace {
struct Kin { double eta; Eigen::Vector3d vel, acc; };
Kin kinematics(const std::vector<double>& omega, const std::vector<double>& amp, const std::vector<double>& phase,
               const std::vector<double>& k, const Eigen::Vector3d& p, double t, double depth, double mwl, bool wheeler) {
    Kin r{0.0, Eigen::Vector3d(0.0, 0.0, 0.0), Eigen::Vector3d(0.0, 0.0, 0.0)};
    const double pos[3] = {p.x(), p.y(), p.z()};
    double v[3], a[3];
    hc_throw_on_error(hc_wave_kinematics(int(omega.size()), omega.data(), amp.data(), phase.data(), k.data(), pos, t, depth,
                                         mwl, wheeler ? 1 : 0, &r.eta, v, a));
    r.vel = Eigen::Vector3d(v[0], v[1], v[2]);
    r.acc = Eigen::Vector3d(a[0], a[1], a[2]);
    return r;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// WaveBase
// ---------------------------------------------------------------------------------------------
void WaveBase::Bind(hc_ensemble* ens, const HydroData::SimulationParameters& sim, unsigned int num_bodies) {
    ens_ = ens;
    bound_bodies_ = num_bodies;
    water_depth_ = sim.water_depth;
    g_ = sim.g;
}

Eigen::VectorXd WaveBase::DeviceForce(double t) const {
    if (!ens_)
        throw std::runtime_error("wave object is not attached to a TestHydro (no device ensemble to evaluate the force)");
    Eigen::VectorXd f(6 * bound_bodies_);
    hc_throw_on_error(hc_waves_force_at_time(ens_, t, f.data()));
    return f;
}

// ---------------------------------------------------------------------------------------------
// NoWave (src/wave_types.cpp:257-264)
// ---------------------------------------------------------------------------------------------
void NoWave::Bind(hc_ensemble* ens, const HydroData::SimulationParameters& sim, unsigned int num_bodies) {
    WaveBase::Bind(ens, sim, num_bodies);
    hc_throw_on_error(hc_waves_none(ens));
}
Eigen::VectorXd NoWave::GetForceAtTime(double) {
    Eigen::VectorXd f(6 * num_bodies_);
    f.setZero();
    return f;
}

// ---------------------------------------------------------------------------------------------
// RegularWave (src/wave_types.cpp:266-352)
// ---------------------------------------------------------------------------------------------
RegularWave::RegularWave() { num_bodies_ = 1; }
RegularWave::RegularWave(unsigned int num_b) { num_bodies_ = num_b; }

void RegularWave::AddH5Data(std::vector<HydroData::RegularWaveInfo>&, HydroData::SimulationParameters& sim_data) {
    // the excitation tables already live in the device tables; only the scalars are kept here
    water_depth_ = sim_data.water_depth;
    g_ = sim_data.g;
}

void RegularWave::Bind(hc_ensemble* ens, const HydroData::SimulationParameters& sim, unsigned int num_bodies) {
    WaveBase::Bind(ens, sim, num_bodies);
    hc_throw_on_error(hc_waves_regular(ens, 1, &regular_wave_amplitude_, &regular_wave_omega_, &regular_wave_phase_));
}

void RegularWave::Initialize() {
    hc_throw_on_error(hc_compute_wave_number(regular_wave_omega_, water_depth_, g_, &wavenumber_));
}

Eigen::VectorXd RegularWave::GetForceAtTime(double t) { return DeviceForce(t); }

Eigen::VectorXd RegularWave::GetExcitationMag() const {
    Eigen::VectorXd m(6 * bound_bodies_);
    hc_throw_on_error(hc_waves_regular_coeffs(ens_, 0, m.data(), nullptr, nullptr));
    return m;
}
Eigen::VectorXd RegularWave::GetExcitationPhase() const {
    Eigen::VectorXd p(6 * bound_bodies_);
    hc_throw_on_error(hc_waves_regular_coeffs(ens_, 0, nullptr, p.data(), nullptr));
    return p;
}

double RegularWave::GetElevation(const Eigen::Vector3d& position, double time) {
    return kinematics({regular_wave_omega_}, {regular_wave_amplitude_}, {regular_wave_phase_}, {wavenumber_}, position, time,
                      water_depth_, mwl_, false).eta;
}
Eigen::Vector3d RegularWave::GetVelocity(const Eigen::Vector3d& position, double time) {
    return kinematics({regular_wave_omega_}, {regular_wave_amplitude_}, {regular_wave_phase_}, {wavenumber_}, position, time,
                      water_depth_, mwl_, false).vel;
}
Eigen::Vector3d RegularWave::GetAcceleration(const Eigen::Vector3d& position, double time) {
    return kinematics({regular_wave_omega_}, {regular_wave_amplitude_}, {regular_wave_phase_}, {wavenumber_}, position, time,
                      water_depth_, mwl_, false).acc;
}

// ---------------------------------------------------------------------------------------------
// IrregularWaves (src/wave_types.cpp:430-570,717-774)
// ---------------------------------------------------------------------------------------------
IrregularWaves::IrregularWaves(const IrregularWaveParams& params) : params_(params) {}

void IrregularWaves::AddH5Data(std::vector<HydroData::IrregularWaveInfo>&, HydroData::SimulationParameters& sim_data) {
    water_depth_ = sim_data.water_depth;
    g_ = sim_data.g;
}

void IrregularWaves::Bind(hc_ensemble* ens, const HydroData::SimulationParameters& sim, unsigned int num_bodies) {
    WaveBase::Bind(ens, sim, num_bodies);
    if (!params_.eta_file_path_.empty()) {
        // InitializeIRFVectors (src/wave_types.cpp:451-453): an eta file replaces the spectrum.  The reference snapshot
        // fills no time grid on this branch; the file's own time column is the grid here.
        std::vector<double> time, eta;
        ReadEtaFromFile(params_.eta_file_path_, time, eta);
        hc_throw_on_error(hc_waves_irregular_series(ens, params_.simulation_dt_, int(time.size()), time.data(), eta.data(), 0));
        spectrum_fetched_ = false;
        comp_omega_.clear(); comp_amp_.clear();
        return;
    }
    hc_irregular_params q;
    hc_irregular_default_params(&q);
    q.simulation_dt = params_.simulation_dt_;
    q.simulation_duration = params_.simulation_duration_;
    q.ramp_duration = params_.ramp_duration_;
    q.wave_height = params_.wave_height_;
    q.wave_period = params_.wave_period_;
    q.frequency_min = params_.frequency_min_;
    q.frequency_max = params_.frequency_max_;
    q.nfrequencies = params_.nfrequencies_;
    q.peak_enhancement_factor = params_.peak_enhancement_factor_;
    q.is_normalized = params_.is_normalized_ ? 1 : 0;
    q.seed = params_.seed_;
    hc_throw_on_error(hc_waves_irregular(ens, &q, nullptr, nullptr, nullptr));
    spectrum_fetched_ = false;
    comp_omega_.clear(); comp_amp_.clear();
}

// "time : eta" per line (src/wave_types.cpp:480-500), same messages
void IrregularWaves::ReadEtaFromFile(const std::string& path, std::vector<double>& time_data, std::vector<double>& eta_data) {
    std::ifstream file(path);
    if (!file) throw std::runtime_error("Unable to open file at: " + path + ".");
    std::string line;
    while (std::getline(file, line)) {
        std::stringstream ss(line);
        double time = 0.0, eta = 0.0;
        char delimiter = 0;
        if (!(ss >> time >> delimiter >> eta) || delimiter != ':') throw std::runtime_error("Could not parse line: " + line + ".");
        time_data.push_back(time);
        eta_data.push_back(eta);
    }
}

void IrregularWaves::FetchSpectrum() const {
    if (spectrum_fetched_) return;
    if (!ens_) throw std::runtime_error("Spectrum has not been created. Initialize with wave height and period to create spectrum.");
    int nf = 0, ne = 0;
    hc_throw_on_error(hc_waves_irregular_sizes(ens_, &nf, &ne, nullptr));
    if (nf == 0) throw std::runtime_error("Spectrum has not been created. Initialize with wave height and period to create spectrum.");
    freqs_.resize(nf); S_.resize(nf); widths_.resize(nf); phases_.resize(nf); wavenumbers_.resize(nf);
    hc_throw_on_error(hc_waves_irregular_spectrum(ens_, 0, freqs_.data(), S_.data(), widths_.data(), phases_.data(),
                                                  wavenumbers_.data()));
    spectrum_fetched_ = true;
}

std::vector<double> IrregularWaves::GetSpectrum() { FetchSpectrum(); return S_; }
std::vector<double> IrregularWaves::GetFrequenciesHz() const { FetchSpectrum(); return freqs_; }

std::vector<double> IrregularWaves::GetFreeSurfaceElevation() {
    int nf = 0, ne = 0;
    if (!ens_) return {};
    hc_throw_on_error(hc_waves_irregular_sizes(ens_, &nf, &ne, nullptr));
    std::vector<double> eta(ne);
    hc_throw_on_error(hc_waves_irregular_eta(ens_, 0, nullptr, eta.data()));
    return eta;
}
std::vector<double> IrregularWaves::GetFreeSurfaceTime() const {
    int nf = 0, ne = 0;
    if (!ens_) return {};
    hc_throw_on_error(hc_waves_irregular_sizes(ens_, &nf, &ne, nullptr));
    std::vector<double> t(ne);
    hc_throw_on_error(hc_waves_irregular_eta(ens_, 0, t.data(), nullptr));
    return t;
}

Eigen::VectorXd IrregularWaves::GetForceAtTime(double t) { return DeviceForce(t); }

// per-component omega = 2 pi f and amplitude = sqrt(2 S df) exactly as GetEtaIrregular forms them (:38-40)
void IrregularWaves::FetchComponents() const {
    int nf = 0, ne = 0;
    if (ens_ && hc_waves_irregular_sizes(ens_, &nf, &ne, nullptr) == HC_OK && nf == 0) {
        // no spectrum (eta imported from a file, or zero wave height): the reference's sums run over empty vectors
        comp_omega_.clear(); comp_amp_.clear(); phases_.clear(); wavenumbers_.clear();
        return;
    }
    FetchSpectrum();
    if (comp_omega_.size() == freqs_.size()) return;
    comp_omega_.resize(freqs_.size()); comp_amp_.resize(freqs_.size());
    for (size_t i = 0; i < freqs_.size(); ++i) {
        comp_amp_[i] = std::sqrt(2 * S_[i] * widths_[i]);
        comp_omega_[i] = 2 * M_PI * freqs_[i];
    }
}

double IrregularWaves::GetElevation(const Eigen::Vector3d& position, double time) {
    FetchComponents();
    return kinematics(comp_omega_, comp_amp_, phases_, wavenumbers_, position, time, water_depth_, mwl_, false).eta;
}

Eigen::Vector3d IrregularWaves::GetVelocity(const Eigen::Vector3d& position, double time) {
    FetchComponents();       // Wheeler stretching (:516-525) inside hc_wave_kinematics
    return kinematics(comp_omega_, comp_amp_, phases_, wavenumbers_, position, time, water_depth_, mwl_,
                      params_.wave_stretching_).vel;
}

Eigen::Vector3d IrregularWaves::GetAcceleration(const Eigen::Vector3d& position, double time) {
    FetchComponents();
    return kinematics(comp_omega_, comp_amp_, phases_, wavenumbers_, position, time, water_depth_, mwl_,
                      params_.wave_stretching_).acc;
}

// Free-surface strip mesh for visualisation: two vertices per elevation sample, two triangles per quad.
void IrregularWaves::SetUpWaveMesh(std::string filename) {
    mesh_file_name_ = filename;
    const std::vector<double> eta = GetFreeSurfaceElevation();
    const int n = static_cast<int>(std::ceil(params_.simulation_duration_ / params_.simulation_dt_)) + 1;
    std::ofstream out(filename);
    if (!out) { std::cerr << "Failed to open " << filename << std::endl; return; }
    out << "# Wavefront OBJ file exported by hydrochrono_b200\n";
    out << std::fixed << std::setprecision(6);
    const int count = std::min<int>(n, int(eta.size()));
    for (int i = 0; i < count; ++i) {
        const double x = -1.0 * (i * params_.simulation_dt_);
        out << "v " << std::setw(14) << x << ' ' << std::setw(14) << -10.0 << ' ' << std::setw(14) << eta[i] << "\n";
        out << "v " << std::setw(14) << x << ' ' << std::setw(14) << 10.0 << ' ' << std::setw(14) << eta[i] << "\n";
    }
    for (int i = 0; i + 1 < count; ++i) {
        out << "f " << 2 * i + 1 << ' ' << 2 * i + 2 << ' ' << 2 * i + 4 << "\n";
        out << "f " << 2 * i + 1 << ' ' << 2 * i + 4 << ' ' << 2 * i + 3 << "\n";
    }
}
std::string IrregularWaves::GetMeshFile() { return mesh_file_name_; }
Eigen::Vector3<double> IrregularWaves::GetWaveMeshVelocity() { return Eigen::Vector3d(1.0, 0, 0); }
