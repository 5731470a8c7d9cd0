// SimulationExporter with the reference's public surface (include/hydroc/simulation_exporter.h): buffers the
// per-step body states and writes one results .h5 in schema v0.3 at Finalize() -- /inputs/model/bodies/<name>/*,
// /inputs/simulation/{time,environment,waves[/irregular]}, /results/time/time,
// /results/model/bodies/<name>/{position,velocity,acceleration,orientation,orientation_xyz,angular_velocity},
// /inputs/model/{joints,tsdas,rsdas}/names + per-link metadata, /results/model/tsdas/<name>/{force_vec,force_mag,
// extension,speed,spring_force,damping_force,reaction_force_body1,2}, /results/model/joints/<name>/reaction{1,2}_{force,
// torque}, /meta -- through the library's libhdf5-free writer (hc_h5_writer_*).  Joint reactions are reported in the
// world frame (the reference reports Chrono's link-frame wrenches); the stand-in system has no rotational dampers.
#ifndef HYDROC_B200_SIMULATION_EXPORTER_H
#define HYDROC_B200_SIMULATION_EXPORTER_H

#include <memory>
#include <string>
#include <vector>

#include <chrono_compat/chrono_compat.h>

namespace hydroc {

enum class H5Verbosity { Quiet = 0, Verbose = 1 };

class SimulationExporter {
  public:
    struct Options {
        std::string output_path;
        std::string model_yaml;
        std::string hydro_yaml;
        std::string input_model_file;
        std::string input_simulation_file;
        std::string input_hydro_file;
        std::string output_directory;
        std::string output_tag;
        std::string setup_yaml_text;
        std::string setup_yaml_path;
        int run_steps = 0;
        double run_dt = 0.0;
        double run_time_final = 0.0;
        std::string run_started_at_utc;
        std::string run_finished_at_utc;
        double run_wall_time_s = 0.0;
        std::string scenario_type;   // still | regular | irregular | no_wave
        double scenario_H = 0.0;
        double scenario_T = 0.0;
        double scenario_Hs = 0.0;
        double scenario_Tp = 0.0;
        int scenario_seed = -1;
        H5Verbosity verbosity = H5Verbosity::Quiet;
    };

    SimulationExporter(const Options& opts);
    ~SimulationExporter() noexcept;
    SimulationExporter(const SimulationExporter&) = delete;
    SimulationExporter& operator=(const SimulationExporter&) = delete;

    void WriteSimulationInfo(chrono::ChSystem* system, const std::string& chrono_version, const std::string& model_name,
                             double timestep, double duration_seconds);
    void WriteModel(chrono::ChSystem* system);
    void BeginResults(chrono::ChSystem* system, int expected_steps);
    void RecordStep(chrono::ChSystem* system);
    void Finalize();
    void WriteIrregularInputs(const std::vector<double>& frequencies_hz, const std::vector<double>& spectral_densities,
                              const std::vector<double>& free_surface_time, const std::vector<double>& free_surface_eta);
    void SetRunMetadata(const std::string& started_at_utc, const std::string& finished_at_utc, double wall_time_s,
                        int steps, double dt_s, double time_final_s);

  private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};

}  // namespace hydroc

#endif
