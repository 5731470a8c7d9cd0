// WaveBase family with the reference's public surface (include/hydroc/wave_types.h:14-435).
//
// The per-step wave force (GetForceAtTime) is computed on the GPU: a wave object is bound to the device
// ensemble when it is attached to a TestHydro (TestHydro::AddWaves).  Kinematics (GetElevation / GetVelocity /
// GetAcceleration) are off the per-step path in the reference as well and are evaluated on the host.
#ifndef HYDROC_B200_WAVE_TYPES_H
#define HYDROC_B200_WAVE_TYPES_H
#pragma once

#include <array>
#include <string>
#include <vector>

#include <hydroc/h5fileinfo.h>

// src/wave_types.cpp:679-715
Eigen::VectorXd PiersonMoskowitzSpectrumHz(Eigen::VectorXd& f, double Hs, double Tp);
Eigen::VectorXd JONSWAPSpectrumHz(Eigen::VectorXd& f, double Hs, double Tp, double gamma = 3.3, bool is_normalized = false);

enum class WaveMode { noWaveCIC = 0, regular = 1, irregular = 2 };

class TestHydro;

class WaveBase {
  public:
    virtual ~WaveBase() = default;
    virtual void Initialize() = 0;
    // 6N-dimensional wave force on the hydro bodies at time t (device evaluation once attached to a TestHydro)
    virtual Eigen::VectorXd GetForceAtTime(double t) = 0;
    virtual WaveMode GetWaveMode() = 0;
    virtual double GetElevation(const Eigen::Vector3d& position, double time) = 0;
    virtual Eigen::Vector3d GetVelocity(const Eigen::Vector3d& position, double time) = 0;
    virtual Eigen::Vector3d GetAcceleration(const Eigen::Vector3d& position, double time) = 0;

    double mwl_ = 0.0;          // mean water level
    double g_ = 9.81;           // gravitational acceleration
    double water_depth_ = 0.0;  // water depth

  protected:
    friend class TestHydro;
    // binds the wave to the device ensemble that evaluates its force; called by TestHydro::AddWaves
    virtual void Bind(hc_ensemble* ens, const HydroData::SimulationParameters& sim, unsigned int num_bodies);
    Eigen::VectorXd DeviceForce(double t) const;
    hc_ensemble* ens_ = nullptr;
    unsigned int bound_bodies_ = 0;
};

class NoWave : public WaveBase {
  public:
    NoWave() { num_bodies_ = 1; }
    NoWave(unsigned int num_b) { num_bodies_ = num_b; }
    void Initialize() override {}
    Eigen::VectorXd GetForceAtTime(double t) override;
    WaveMode GetWaveMode() override { return mode_; }
    double GetElevation(const Eigen::Vector3d&, double) override { return 0.0; }
    Eigen::Vector3d GetVelocity(const Eigen::Vector3d&, double) override { return Eigen::Vector3d(0.0, 0.0, 0.0); }
    Eigen::Vector3d GetAcceleration(const Eigen::Vector3d&, double) override { return Eigen::Vector3d(0.0, 0.0, 0.0); }

  protected:
    void Bind(hc_ensemble* ens, const HydroData::SimulationParameters& sim, unsigned int num_bodies) override;

  private:
    unsigned int num_bodies_;
    const WaveMode mode_ = WaveMode::noWaveCIC;
};

class RegularWave : public WaveBase {
  public:
    RegularWave();
    RegularWave(unsigned int num_b);
    void Initialize() override;
    Eigen::VectorXd GetForceAtTime(double t) override;
    WaveMode GetWaveMode() override { return mode_; }

    // user input
    double regular_wave_amplitude_ = 0.0;
    double regular_wave_omega_ = 0.0;
    double regular_wave_phase_ = 0.0;

    void AddH5Data(std::vector<HydroData::RegularWaveInfo>& reg_h5_data, HydroData::SimulationParameters& sim_data);
    double GetElevation(const Eigen::Vector3d& position, double time) override;
    Eigen::Vector3d GetVelocity(const Eigen::Vector3d& position, double time) override;
    Eigen::Vector3d GetAcceleration(const Eigen::Vector3d& position, double time) override;

    // interpolated excitation coefficients (excitation_force_mag_ / _phase_ in the reference), after attach
    Eigen::VectorXd GetExcitationMag() const;
    Eigen::VectorXd GetExcitationPhase() const;

  protected:
    void Bind(hc_ensemble* ens, const HydroData::SimulationParameters& sim, unsigned int num_bodies) override;

  private:
    unsigned int num_bodies_;
    const WaveMode mode_ = WaveMode::regular;
    double wavenumber_ = 0.0;
};

struct IrregularWaveParams {   // include/hydroc/wave_types.h:277-292
    unsigned int num_bodies_ = 1;
    double simulation_dt_ = 0.0;
    double simulation_duration_ = 0.0;
    double ramp_duration_ = 0.0;
    std::string eta_file_path_;
    double wave_height_ = 0.0;
    double wave_period_ = 0.0;
    double frequency_min_ = 0.001;
    double frequency_max_ = 1.0;
    double nfrequencies_ = 0;
    double peak_enhancement_factor_ = 1.0;
    bool is_normalized_ = false;
    int seed_ = 1;
    bool wave_stretching_ = true;
};

class IrregularWaves : public WaveBase {
  public:
    IrregularWaves(const IrregularWaveParams& params);
    void Initialize() override {}

    std::vector<double> GetSpectrum();               // S(f); throws if no spectrum was created (:461-467)
    std::vector<double> GetFreeSurfaceElevation();   // precomputed eta(t) samples (device -> host)
    std::vector<double> GetFreeSurfaceTime() const;
    std::vector<double> GetFrequenciesHz() const;
    Eigen::VectorXd GetForceAtTime(double t) override;
    WaveMode GetWaveMode() override { return mode_; }

    void AddH5Data(std::vector<HydroData::IrregularWaveInfo>& irreg_h5_data, HydroData::SimulationParameters& sim_data);
    double GetElevation(const Eigen::Vector3d& position, double time) override;
    Eigen::Vector3d GetVelocity(const Eigen::Vector3d& position, double time) override;
    Eigen::Vector3d GetAcceleration(const Eigen::Vector3d& position, double time) override;

    const IrregularWaveParams& GetParams() const { return params_; }

    // visualisation helpers of the reference (free-surface OBJ mesh); kept for source compatibility
    void SetUpWaveMesh(std::string filename = "fse_mesh.obj");
    std::string GetMeshFile();
    Eigen::Vector3<double> GetWaveMeshVelocity();

  protected:
    void Bind(hc_ensemble* ens, const HydroData::SimulationParameters& sim, unsigned int num_bodies) override;

  private:
    void FetchSpectrum() const;
    /// "time : eta" lines of IrregularWaveParams::eta_file_path_ (reference: src/wave_types.cpp:480-500).
    static void ReadEtaFromFile(const std::string& path, std::vector<double>& time_data, std::vector<double>& eta_data);
    void FetchComponents() const;
    IrregularWaveParams params_;
    const WaveMode mode_ = WaveMode::irregular;
    mutable std::vector<double> freqs_, S_, widths_, phases_, wavenumbers_;
    mutable std::vector<double> comp_omega_, comp_amp_;
    mutable bool spectrum_fetched_ = false;
    std::string mesh_file_name_;
};

#endif
