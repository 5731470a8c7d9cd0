// TestHydro / ForceFunc6d / ComponentFunc with the reference's public surface
// (include/hydroc/hydro_forces.h:45-348).  The force computation itself lives in libhydrochrono_b200
// (CUDA, sm_100a); this layer gathers the Chrono body state, calls hc_step and serves Chrono's ChFunction hooks.
#ifndef HYDROC_B200_HYDRO_FORCES_H
#define HYDROC_B200_HYDRO_FORCES_H
#pragma once

#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include <chrono_compat/chrono_compat.h>
#include <hydroc/h5fileinfo.h>
#include <hydroc/wave_types.h>

using namespace chrono;

class ForceFunc6d;
class TestHydro;
class ChLoadAddedMass;

class ComponentFunc : public ChFunction {
  public:
    ComponentFunc();
    ComponentFunc(const ComponentFunc& old);
    ComponentFunc(ForceFunc6d* b, int i);
    virtual ComponentFunc* Clone() const override;
    // x is the simulation time Chrono passes in; the force of DoF index_ on the body is returned
    virtual double GetVal(double x) const override;

  private:
    ForceFunc6d* base_;
    int index_;
};

class ForceFunc6d {
  public:
    ForceFunc6d();
    ForceFunc6d(std::shared_ptr<ChBody> object, TestHydro* all_hydro_forces_user);
    ForceFunc6d(const ForceFunc6d& old);
    double CoordinateFunc(int i);

  private:
    void SetForce();
    void SetTorque();
    void ApplyForceAndTorqueToBody();

    std::shared_ptr<ChBody> body_;
    int b_num_;   // 1-indexed, parsed from the body name "bodyN"
    ComponentFunc forces_[6];
    std::shared_ptr<ComponentFunc> force_ptrs_[6];
    std::shared_ptr<ChForce> chrono_force_;
    std::shared_ptr<ChForce> chrono_torque_;
    TestHydro* all_hydro_forces_;
};

struct HydroProfileStats {
    double hydrostatics_seconds = 0.0;
    double radiation_seconds = 0.0;
    double waves_seconds = 0.0;
    int hydrostatics_calls = 0;
    int radiation_calls = 0;
    int waves_calls = 0;
};

class TestHydro {
  public:
    TestHydro() = delete;
    TestHydro(std::vector<std::shared_ptr<ChBody>> user_bodies, std::string h5_file_name,
              std::shared_ptr<WaveBase> waves = std::make_shared<NoWave>());
    // same, from tables that are already in memory (takes ownership of the handle)
    TestHydro(std::vector<std::shared_ptr<ChBody>> user_bodies, hc_tables* tables,
              std::shared_ptr<WaveBase> waves = std::make_shared<NoWave>());
    TestHydro(const TestHydro& old) = delete;
    TestHydro& operator=(const TestHydro& rhs) = delete;
    ~TestHydro();

    void AddWaves(std::shared_ptr<WaveBase> waves);

    // Components of the force at the current Chrono time.  As in the reference, each evaluates at most once per
    // time value: the first call at a new time gathers the body state and runs the device step.
    std::vector<double> ComputeForceHydrostatics();
    std::vector<double> ComputeForceRadiationDampingConv();
    Eigen::VectorXd ComputeForceWaves();
    std::shared_ptr<WaveBase> GetWave() const { return user_waves_; }

    double GetRIRFval(int row, int col, int st);

    enum class RadiationConvolutionMode { Baseline, TaperedDirect };
    void SetRadiationConvolutionMode(RadiationConvolutionMode mode);

    struct TaperedDirectOptions {
        std::string smoothing = "sg";
        int window_length = 5;
        double rirf_end_time = -1.0;
        double taper_start_percent = 0.8;
        double taper_end_percent = 1.0;
        double taper_final_amplitude = 0.0;
        bool export_plot_csv = false;
    };
    void SetTaperedDirectOptions(const TaperedDirectOptions& opts);
    void SetDiagnosticsOutputDirectory(const std::string& dir) { diagnostics_output_dir_ = dir; }

    // total force on body b (1-based) in DoF i; cached per time value (src/hydro_forces.cpp:727-767)
    double CoordinateFuncForBody(int b, int i);

    HydroProfileStats GetProfileStats() const;

    // --- additions of this implementation ---
    hc_ensemble* ensemble() const { return ens_; }
    HydroData& GetHydroData() { return file_info_; }

  private:
    void Construct(std::shared_ptr<WaveBase> waves);
    void EvaluateAtCurrentTime();
    void ApplyConvolutionMode();

    std::vector<std::shared_ptr<ChBody>> bodies_;
    int num_bodies_;
    HydroData file_info_;
    std::vector<ForceFunc6d> force_per_body_;
    std::shared_ptr<WaveBase> user_waves_;

    std::vector<double> force_hydrostatic_, force_radiation_damping_, force_waves_, total_force_;
    double prev_time;
    bool components_fetched_ = false;
    std::unique_ptr<std::ofstream> trace_;   // HYDROC_STATE_TRACE=<file>: log of every evaluation (debugging / parity tests)

    std::shared_ptr<ChLoadContainer> my_loadcontainer;
    std::shared_ptr<ChLoadAddedMass> my_loadbodyinertia;

    RadiationConvolutionMode convolution_mode_ = RadiationConvolutionMode::Baseline;
    bool convolution_dirty_ = false;
    TaperedDirectOptions tapered_opts_;
    std::string diagnostics_output_dir_;

    hc_ensemble* ens_ = nullptr;
};

#endif
