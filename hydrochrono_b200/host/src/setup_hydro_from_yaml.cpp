// YAML settings -> wave object + TestHydro (reference src/setup_hydro_from_yaml.cpp:28-193).
#include <hydroc/setup_hydro_from_yaml.h>

#include <algorithm>
#include <cmath>
#include <iostream>
#include <stdexcept>

#include <hydroc/hydro_forces.h>
#include <hydroc/wave_types.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

std::string lowercase(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    return s;
}

// height -> amplitude (H/2), period -> omega (2 pi / T); irregular: seed defaults to 1; YAML cannot select gamma,
// so irregular waves run with the IrregularWaveParams default (gamma = 1, Pierson-Moskowitz)
std::shared_ptr<WaveBase> MakeWave(const WaveSettings& ws, unsigned int num_bodies, double timestep, double sim_duration,
                                   double ramp_duration) {
    const std::string type = lowercase(ws.type);
    if (type == "regular") {
        auto w = std::make_shared<RegularWave>(num_bodies);
        w->regular_wave_amplitude_ = ws.height / 2.0;
        w->regular_wave_omega_ = 2.0 * M_PI / ws.period;
        w->regular_wave_phase_ = ws.phase;
        return w;
    }
    if (type == "irregular") {
        IrregularWaveParams p;
        p.num_bodies_ = num_bodies;
        p.simulation_dt_ = timestep;
        p.simulation_duration_ = sim_duration;
        p.ramp_duration_ = ramp_duration;
        p.wave_height_ = ws.height;
        p.wave_period_ = ws.period;
        p.seed_ = (ws.seed > 0 ? ws.seed : 1);
        return std::make_shared<IrregularWaves>(p);
    }
    if (type == "no_wave" || type == "still_ci" || type == "still") return std::make_shared<NoWave>(num_bodies);
    throw std::runtime_error("Unsupported wave type: " + ws.type);
}

}  // namespace

std::unique_ptr<TestHydro> SetupHydroFromYAML(const YAMLHydroData& hydro_data,
                                              const std::vector<std::shared_ptr<chrono::ChBody>>& bodies, double timestep,
                                              double sim_duration, double ramp_duration) {
    // match hydro bodies to Chrono bodies by name; the first body's h5_file is used for all (reference :91-95)
    std::string h5_file_path;
    if (!hydro_data.bodies.empty()) h5_file_path = hydro_data.bodies[0].h5_file;
    std::vector<std::shared_ptr<chrono::ChBody>> matched;
    for (const auto& hb : hydro_data.bodies) {
        bool found = false;
        for (const auto& cb : bodies)
            if (cb->GetName() == hb.name) { matched.push_back(cb); found = true; break; }
        if (!found) std::cerr << "WARNING: Hydrodynamic body '" << hb.name << "' not found in Chrono system" << std::endl;
    }
    if (matched.empty()) throw std::runtime_error("No hydrodynamic bodies found in Chrono system");

    auto wave = MakeWave(hydro_data.waves, static_cast<unsigned int>(matched.size()), timestep, sim_duration, ramp_duration);
    auto hydro = std::make_unique<TestHydro>(matched, h5_file_path, wave);

    if (lowercase(hydro_data.radiation_convolution_mode) == "tapereddirect") {
        hydro->SetRadiationConvolutionMode(TestHydro::RadiationConvolutionMode::TaperedDirect);
        TestHydro::TaperedDirectOptions o;
        o.smoothing = !hydro_data.td_smoothing.empty() ? hydro_data.td_smoothing : o.smoothing;
        o.window_length = std::max(3, hydro_data.td_window_length != 0 ? hydro_data.td_window_length : o.window_length);
        if (o.window_length % 2 == 0) o.window_length += 1;   // enforce odd
        o.rirf_end_time = hydro_data.td_rirf_end_time;
        o.taper_start_percent = hydro_data.td_taper_start_percent;
        o.taper_end_percent = hydro_data.td_taper_end_percent;
        o.taper_final_amplitude = hydro_data.td_taper_final_amplitude;
        o.export_plot_csv = hydro_data.td_export_plot_csv;
        hydro->SetTaperedDirectOptions(o);
    } else {
        hydro->SetRadiationConvolutionMode(TestHydro::RadiationConvolutionMode::Baseline);
    }
    return hydro;
}
