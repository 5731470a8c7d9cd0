// Sphere in regular / irregular waves through the reference's API surface
// (cf. tests/regression/sphere/reg_waves/sphere_reg_waves_test.cpp and irreg_waves/sphere_irreg_waves_test.cpp):
// ground + prismatic joint (heave only) + TSDA damper, RegularWave or IrregularWaves attached to TestHydro.
// usage: demo_sphere_waves <sphere.h5> <out.txt> regular <wave_num 1..10> [duration]
//        demo_sphere_waves <sphere.h5> <out.txt> irregular [duration [eta_dump.txt]]   (dump: "time : eta" lines)
//        demo_sphere_waves <sphere.h5> <out.txt> eta <eta_file.txt> [duration]          (IrregularWaveParams::eta_file_path_)
#include <hydroc/hydro_forces.h>

#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "common_demo.h"

using namespace chrono;

int main(int argc, char* argv[]) {
    if (argc < 4) { std::cerr << "usage: demo_sphere_waves <sphere.h5> <out.txt> regular N | irregular [duration]" << std::endl; return 2; }
    const std::string h5fname = argv[1];
    const bool regular = std::strcmp(argv[3], "regular") == 0;
    const int wave_num = regular && argc > 4 ? std::atoi(argv[4]) : 1;
    double simulationDuration = 600.0;
    if (regular && argc > 5) simulationDuration = std::atof(argv[5]);
    const bool from_file = std::strcmp(argv[3], "eta") == 0;
    if (from_file && argc < 5) { std::cerr << "eta mode needs a file" << std::endl; return 2; }
    if (!regular && !from_file && argc > 4) simulationDuration = std::atof(argv[4]);
    if (from_file && argc > 5) simulationDuration = std::atof(argv[5]);
    const char* eta_dump = (!regular && !from_file && argc > 5) ? argv[5] : nullptr;

    const double task10_wave_amps[] = {0.177, 0.314, 0.380, 0.491, 0.706, 0.961, 1.256, 1.589, 1.962, 2.374};
    const double task10_wave_omegas[] = {2.094395102, 1.570796327, 1.427996661, 1.256637061, 1.047197551,
                                         0.897597901, 0.785398163, 0.698131701, 0.628318531, 0.571198664};
    const double task10_damping_coeffs[] = {398736.034, 118149.758, 90080.857,  161048.558, 322292.419,
                                            479668.979, 633979.761, 784083.286, 932117.647, 1077123.445};

    ChSystemNSC system;
    system.SetGravitationalAcceleration(ChVector3d(0.0, 0.0, -9.81));
    double timestep = 0.015;

    auto ground = chrono_types::make_shared<ChBody>();
    system.AddBody(ground);
    ground->SetPos(ChVector3d(0, 0, -5));
    ground->SetFixed(true);

    std::shared_ptr<ChBody> sphereBody = chrono_types::make_shared<ChBodyEasyMesh>("oes_task10_sphere.obj", 1000, false, true, false);
    system.Add(sphereBody);
    sphereBody->SetName("body1");
    sphereBody->SetPos(ChVector3d(0, 0, -2));
    sphereBody->SetMass(261.8e3);

    auto prismatic = chrono_types::make_shared<ChLinkLockPrismatic>();
    prismatic->Initialize(sphereBody, ground, false, ChFramed(ChVector3d(0, 0, -2)), ChFramed(ChVector3d(0, 0, -5)));
    system.AddLink(prismatic);

    auto spring_1 = chrono_types::make_shared<ChLinkTSDA>();
    spring_1->Initialize(sphereBody, ground, false, ChVector3d(0, 0, -2), ChVector3d(0, 0, -5));
    spring_1->SetSpringCoefficient(0.0);
    spring_1->SetDampingCoefficient(regular ? task10_damping_coeffs[wave_num - 1] : 0.0);
    system.AddLink(spring_1);

    std::vector<std::shared_ptr<ChBody>> bodies;
    bodies.push_back(sphereBody);
    std::vector<double> time_vector, heave_position;
    try {
        std::shared_ptr<WaveBase> waves;
        if (regular) {
            auto my_hydro_inputs = std::make_shared<RegularWave>(1);
            my_hydro_inputs->regular_wave_amplitude_ = task10_wave_amps[wave_num - 1];
            my_hydro_inputs->regular_wave_omega_ = task10_wave_omegas[wave_num - 1];
            waves = my_hydro_inputs;
        } else {
            IrregularWaveParams wave_inputs;
            wave_inputs.num_bodies_ = bodies.size();
            wave_inputs.simulation_dt_ = timestep;
            wave_inputs.simulation_duration_ = 600.0;
            wave_inputs.ramp_duration_ = 60.0;
            wave_inputs.wave_height_ = 2.0;
            wave_inputs.wave_period_ = 12.0;
            wave_inputs.frequency_min_ = 0.001;
            wave_inputs.frequency_max_ = 1.0;
            wave_inputs.nfrequencies_ = 1000;
            if (from_file) wave_inputs.eta_file_path_ = argv[4];      // the series replaces the spectrum (wave_types.cpp:451-453)
            waves = std::make_shared<IrregularWaves>(wave_inputs);
        }
        TestHydro hydro_forces(bodies, h5fname);
        hydro_forces.AddWaves(waves);
        if (eta_dump) {
            auto irr = std::static_pointer_cast<IrregularWaves>(waves);
            const std::vector<double> tt = irr->GetFreeSurfaceTime(), ee = irr->GetFreeSurfaceElevation();
            FILE* f = std::fopen(eta_dump, "w");
            if (!f) { std::cerr << "cannot write " << eta_dump << std::endl; return 1; }
            for (size_t i = 0; i < tt.size(); ++i) std::fprintf(f, "%.17g : %.17g\n", tt[i], ee[i]);
            std::fclose(f);
        }

        while (system.GetChTime() <= simulationDuration) {
            system.DoStepDynamics(timestep);
            time_vector.push_back(system.GetChTime());
            heave_position.push_back(sphereBody->GetPos().z());
        }
        // the WaveBase surface stays usable: force at an arbitrary time, elevation at the origin
        Eigen::VectorXd f = waves->GetForceAtTime(1.0);
        std::cout << "wave heave force at t=1: " << f[2] << " elevation(0,0,0; t=1): "
                  << waves->GetElevation(Eigen::Vector3d(0, 0, 0), 1.0) << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    if (write_heave(argv[2], time_vector, heave_position)) return 1;
    std::cout << "Simulation finished." << std::endl;
    return 0;
}
