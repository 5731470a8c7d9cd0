// Test infrastructure (not product): prints every field of the YAMLHydroData that ReadHydroYAML returns for each file on
// the command line, in a canonical text form.  Compiled twice with the same source:
//   * against the REFERENCE's own parser where it lies (/root/reference/src/hydro_yaml_parser.cpp + hydro_types.h -- the one
//     piece of the reference that builds from its own sources) -> oracle/_ref/ref_yaml_dump
//   * against this repo's parser (hydrochrono_b200/host/src/hydro_yaml_parser.cpp)   -> oracle/_ref/our_yaml_dump
// tests/test_yaml_vs_reference.py diffs the two outputs over a corpus of hydro.yaml files.
#include <cstdio>
#include <exception>
#include <string>

#include "hydro_yaml_parser.h"

static void dump(const YAMLHydroData& d) {
    std::printf("bodies %zu\n", d.bodies.size());
    for (const HydroBody& b : d.bodies) {
        std::printf(" body name=[%s] h5=[%s] exc=%d rad=%d calc=[%s] mode=[%s] smoothing=[%s] window=%d rms=%.17g frac=%.17g csv=%d\n",
                    b.name.c_str(), b.h5_file.c_str(), int(b.include_excitation), int(b.include_radiation),
                    b.radiation_calculation.c_str(), b.radiation_convolution_mode.c_str(), b.td_smoothing.c_str(),
                    b.td_window_length, b.td_rms_threshold_factor, b.td_taper_fraction_remaining, int(b.td_export_plot_csv));
    }
    const WaveSettings& w = d.waves;
    std::printf("waves type=[%s] height=%.17g period=%.17g direction=%.17g phase=%.17g spectrum=[%s] seed=%d\n", w.type.c_str(),
                w.height, w.period, w.direction, w.phase, w.spectrum.c_str(), w.seed);
    std::printf("period_values %zu:", w.period_values.size());
    for (double v : w.period_values) std::printf(" %.17g", v);
    std::printf("\n");
    std::printf("system mode=[%s] smoothing=[%s] window=%d rirf_end=%.17g taper_start=%.17g taper_end=%.17g final=%.17g csv=%d\n",
                d.radiation_convolution_mode.c_str(), d.td_smoothing.c_str(), d.td_window_length, d.td_rirf_end_time,
                d.td_taper_start_percent, d.td_taper_end_percent, d.td_taper_final_amplitude, int(d.td_export_plot_csv));
}

int main(int argc, char* argv[]) {
    for (int i = 1; i < argc; ++i) {
        std::printf("== %s\n", argv[i]);
        try {
            dump(ReadHydroYAML(argv[i]));
        } catch (const std::exception& e) {
            std::printf("EXCEPTION\n");
            std::fprintf(stderr, "%s: %s\n", argv[i], e.what());
        }
    }
    return 0;
}
