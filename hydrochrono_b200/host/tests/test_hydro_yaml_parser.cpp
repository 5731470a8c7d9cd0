// hydro.yaml parser tests, after the reference's tests/unit/test_hydro_yaml_parser.cpp (its tests/data inputs are
// not in the snapshot; the fixtures under tests/data here are recreated from its assertions) plus the sweep forms,
// shorthands, the convolution section and the error paths of src/hydro_yaml_parser.cpp.
#include <hydroc/hydro_yaml_parser.h>

#include <cmath>
#include <filesystem>
#include <fstream>
#include <iostream>

static int failures = 0;
#define CHECK(cond)                                                                          \
    do {                                                                                     \
        if (!(cond)) { std::cerr << "CHECK failed: " #cond " at line " << __LINE__ << std::endl; ++failures; } \
    } while (0)
#define CHECK_NEAR(a, b) CHECK(std::abs((a) - (b)) <= 1e-10)

static std::string g_data;

static std::string write_tmp(const std::string& name, const std::string& text) {
    std::string p = (std::filesystem::temp_directory_path() / name).string();
    std::ofstream f(p);
    f << text;
    return p;
}
template <class F>
static std::string error_of(F&& f) {
    try { f(); } catch (const std::runtime_error& e) { return e.what(); }
    return "";
}

int main(int argc, char* argv[]) {
    g_data = argc > 1 ? argv[1] : "data";

    {   // TestParsesSphereFile + TestResolvesRelativePaths
        YAMLHydroData d = ReadHydroYAML(g_data + "/test_sphere.hydro.yaml");
        CHECK(d.bodies.size() == 1);
        CHECK(d.bodies[0].name == "sphere");
        CHECK(d.bodies[0].h5_file.find("test_sphere.h5") != std::string::npos);
        CHECK(std::filesystem::path(d.bodies[0].h5_file).is_absolute());
        CHECK(d.bodies[0].h5_file.find("hydroData") != std::string::npos);
        CHECK(d.bodies[0].include_excitation && d.bodies[0].include_radiation);
        CHECK(d.bodies[0].radiation_calculation == "convolution");
        CHECK(d.waves.type == "regular");
        CHECK_NEAR(d.waves.height, 1.5);
        CHECK_NEAR(d.waves.period, 7.0);
        CHECK(d.waves.period_values.size() == 1 && d.waves.period_values[0] == 7.0);
        CHECK_NEAR(d.waves.direction, 0.0);
        CHECK_NEAR(d.waves.phase, 0.0);
        CHECK(d.waves.spectrum == "pierson_moskowitz");
        CHECK(d.waves.seed == -1);
        CHECK(d.radiation_convolution_mode == "Baseline");
    }
    {   // TestParsesMultiBodyFile
        YAMLHydroData d = ReadHydroYAML(g_data + "/test_multi.hydro.yaml");
        CHECK(d.bodies.size() == 2);
        CHECK(d.bodies[0].name == "float" && d.bodies[1].name == "spar");
        CHECK(d.bodies[0].h5_file.find("rm3_float.h5") != std::string::npos);
        CHECK(d.bodies[1].h5_file.find("rm3_spar.h5") != std::string::npos);
        CHECK(d.waves.type == "still_ci");
        CHECK_NEAR(d.waves.height, 0.0);
        CHECK_NEAR(d.waves.period, 0.0);
        CHECK(d.waves.period_values.empty());
    }
    {   // sweep (range), shorthand amplitude, convolution section, unknown trailing section ignored
        YAMLHydroData d = ReadHydroYAML(g_data + "/test_sweep_tapered.hydro.yaml");
        CHECK(d.bodies.size() == 2);
        CHECK(!d.bodies[0].include_excitation && d.bodies[1].include_excitation);
        CHECK(d.waves.type == "Regular");
        CHECK_NEAR(d.waves.height, 0.01);                       // 2 * amplitude
        CHECK(d.waves.period_values.size() == 4);
        CHECK_NEAR(d.waves.period_values[0], 10.0);
        CHECK_NEAR(d.waves.period_values[3], 13.0);
        CHECK_NEAR(d.waves.period, 10.0);
        CHECK_NEAR(d.waves.direction, 30.0);
        CHECK_NEAR(d.waves.phase, 0.25);
        CHECK(d.waves.seed == 7);
        CHECK(d.radiation_convolution_mode == "TaperedDirect");
        CHECK(d.td_smoothing == "savitzky_golay");
        CHECK(d.td_window_length == 7);
        CHECK_NEAR(d.td_taper_start_percent, 0.7);
        CHECK_NEAR(d.td_taper_end_percent, 0.95);
        CHECK_NEAR(d.td_taper_final_amplitude, 0.1);
        CHECK_NEAR(d.td_rirf_end_time, 12.5);
        CHECK(d.td_export_plot_csv);
    }
    {   // linspace / values forms, tp shorthand
        std::string p = write_tmp("hc_lin.hydro.yaml",
            "hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: a.h5\n  waves:\n    type: regular\n    h: 2.0\n"
            "    period:\n      linspace: { start: 2.0, stop: 5.0, num: 4 }\n");
        YAMLHydroData d = ReadHydroYAML(p);
        CHECK(d.waves.period_values.size() == 4);
        CHECK_NEAR(d.waves.period_values[1], 3.0);
        CHECK_NEAR(d.waves.height, 2.0);
        p = write_tmp("hc_val.hydro.yaml",
            "hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: a.h5\n  waves:\n    type: regular\n    height: 2.0\n"
            "    period:\n      values: [6.0, 7.5, 9]\n");
        d = ReadHydroYAML(p);
        CHECK(d.waves.period_values.size() == 3);
        CHECK_NEAR(d.waves.period_values[1], 7.5);
        p = write_tmp("hc_tp.hydro.yaml",
            "hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: a.h5\n  waves:\n    type: irregular\n    height: 2.0\n    tp: 12\n    seed: 3\n");
        d = ReadHydroYAML(p);
        CHECK(d.waves.type == "irregular");
        CHECK_NEAR(d.waves.period, 12.0);
        CHECK(d.waves.seed == 3);
    }
    {   // TestHandlesMissingFile
        std::string e = error_of([&] { ReadHydroYAML(g_data + "/nonexistent.hydro.yaml"); });
        CHECK(e.find("Could not open hydro file") != std::string::npos);
        CHECK(e.find("nonexistent.hydro.yaml") != std::string::npos);
    }
    {   // TestHandlesMalformedYAML
        std::string p = write_tmp("hc_malformed.hydro.yaml", "bodies:\n  - name: test\n    h5_file: test.h5\n");
        std::string e = error_of([&] { ReadHydroYAML(p); });
        CHECK(e.find("No 'hydrodynamics:' section found") != std::string::npos);
    }
    {   // validation errors of the parser
        std::string p = write_tmp("hc_noheight.hydro.yaml", "hydrodynamics:\n  bodies:\n    - name: test\n  waves:\n    type: regular\n");
        CHECK(error_of([&] { ReadHydroYAML(p); }).find("regular requires wave height") != std::string::npos);
        p = write_tmp("hc_incons.hydro.yaml", "hydrodynamics:\n  waves:\n    type: regular\n    height: 2.0\n    amplitude: 0.5\n    period: 8\n");
        CHECK(error_of([&] { ReadHydroYAML(p); }).find("inconsistent") != std::string::npos);
        p = write_tmp("hc_badlin.hydro.yaml", "hydrodynamics:\n  waves:\n    type: regular\n    height: 2.0\n    period:\n      linspace: { start: 2.0, stop: 5.0, num: 1 }\n");
        CHECK(error_of([&] { ReadHydroYAML(p); }).find("invalid linspace") != std::string::npos);
        p = write_tmp("hc_badrange.hydro.yaml", "hydrodynamics:\n  waves:\n    type: regular\n    height: 2.0\n    period:\n      range: { start: 9.0, stop: 5.0, step: 1.0 }\n");
        CHECK(error_of([&] { ReadHydroYAML(p); }).find("invalid range") != std::string::npos);
        p = write_tmp("hc_two.hydro.yaml", "hydrodynamics:\n  waves:\n    type: regular\n    height: 2.0\n    period:\n      values: [1, 2]\n      linspace: { start: 2.0, stop: 5.0, num: 3 }\n");
        CHECK(error_of([&] { ReadHydroYAML(p); }).find("multiple forms") != std::string::npos);
    }
    {   // defaults for optional fields (still water: no validation of height/period)
        std::string p = write_tmp("hc_min.hydro.yaml", "hydrodynamics:\n  bodies:\n    - name: test\n      h5_file: test.h5\n  waves:\n    type: still\n");
        YAMLHydroData d = ReadHydroYAML(p);
        CHECK(d.bodies.size() == 1 && d.bodies[0].name == "test");
        CHECK(d.bodies[0].include_excitation && d.bodies[0].include_radiation);
        CHECK_NEAR(d.waves.height, 0.0);
        CHECK(d.waves.spectrum == "pierson_moskowitz");
        CHECK(d.td_window_length == 5 && d.td_rirf_end_time == -1.0);
    }
    if (failures) { std::cerr << failures << " check(s) failed" << std::endl; return 1; }
    std::cout << "All hydro YAML parser tests passed" << std::endl;
    return 0;
}
