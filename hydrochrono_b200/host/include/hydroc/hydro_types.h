// Parsed hydro.yaml content (reference src/hydro_types.h).
#ifndef HYDROC_B200_HYDRO_TYPES_H
#define HYDROC_B200_HYDRO_TYPES_H

#include <string>
#include <vector>

struct HydroBody {
    std::string name = "";
    std::string h5_file = "";
    bool include_excitation = true;
    bool include_radiation = true;
    std::string radiation_calculation = "convolution";
    std::string radiation_convolution_mode = "Baseline";
    std::string td_smoothing = "sg";
    int td_window_length = 5;
    double td_rms_threshold_factor = 0.02;
    double td_taper_fraction_remaining = 0.25;
    bool td_export_plot_csv = false;
};

struct WaveSettings {
    std::string type = "regular";   // "regular", "irregular", "no_wave" / "still" / "still_ci"
    double height = 0.0;
    double period = 0.0;
    double direction = 0.0;
    double phase = 0.0;
    std::string spectrum = "pierson_moskowitz";
    int seed = -1;                  // -1: unset
    std::vector<double> period_values;   // expanded sweep; mirrors `period` when a scalar was given
};

struct YAMLHydroData {
    std::vector<HydroBody> bodies;
    WaveSettings waves;
    std::string radiation_convolution_mode = "Baseline";   // Baseline | TaperedDirect
    std::string td_smoothing = "sg";
    int td_window_length = 5;
    double td_rirf_end_time = -1.0;
    double td_taper_start_percent = 0.8;
    double td_taper_end_percent = 1.0;
    double td_taper_final_amplitude = 0.0;
    bool td_export_plot_csv = false;
};

#endif
