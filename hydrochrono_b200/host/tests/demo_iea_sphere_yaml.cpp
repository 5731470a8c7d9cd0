// The reference's CLI regression case iea_sphere/decay (tests/regression/run_hydrochrono/iea_sphere/decay/inputs):
// body1 at z = -1 m, mass 261800 kg, prismatic heave joint to ground, g = 9.8, dt = 0.01, 40 s, hydro set up from a
// hydro.yaml (waves: still) via ReadHydroYAML + SetupHydroFromYAML, results exported to HDF5 in schema v0.3.
// usage: demo_iea_sphere_yaml <hydro.yaml> <results.h5> [irregular]
#include <hydroc/hydro_forces.h>
#include <hydroc/hydro_yaml_parser.h>
#include <hydroc/setup_hydro_from_yaml.h>
#include <hydroc/simulation_exporter.h>

#include <iostream>

using namespace chrono;

int main(int argc, char* argv[]) {
    if (argc < 3) { std::cerr << "usage: demo_iea_sphere_yaml <hydro.yaml> <results.h5>" << std::endl; return 2; }
    const double timestep = 0.01, end_time = 40.0;
    try {
        ChSystemSMC system;
        system.SetGravitationalAcceleration(ChVector3d(0, 0, -9.8));
        system.SetTimestepperType(ChTimestepper::Type::HHT);   // the stand-in system steps with linearised Euler

        auto body1 = chrono_types::make_shared<ChBody>();
        body1->SetName("body1");
        body1->SetPos(ChVector3d(0, 0, -1.0));
        body1->SetMass(261800);
        body1->SetInertiaXX(ChVector3d(999, 999, 999));
        system.AddBody(body1);
        auto ground = chrono_types::make_shared<ChBody>();
        ground->SetName("ground");
        ground->SetMass(999);
        ground->SetFixed(true);
        system.AddBody(ground);
        auto joint = chrono_types::make_shared<ChLinkLockPrismatic>();
        joint->Initialize(ground, body1, ChFramed(ChVector3d(0, 0, 0)));
        system.AddLink(joint);

        YAMLHydroData hydro_data = ReadHydroYAML(argv[1]);
        std::unique_ptr<TestHydro> hydro = SetupHydroFromYAML(hydro_data, system.GetBodies(), timestep, end_time, 0.0);

        hydroc::SimulationExporter::Options opts;
        opts.output_path = argv[2];
        opts.input_hydro_file = argv[1];
        opts.scenario_type = hydro_data.waves.type;
        opts.scenario_Hs = hydro_data.waves.height;
        opts.scenario_Tp = hydro_data.waves.period;
        opts.scenario_seed = hydro_data.waves.seed;
        hydroc::SimulationExporter exporter(opts);
        exporter.WriteSimulationInfo(&system, CHRONO_VERSION, "iea_sphere_model", timestep, end_time);
        exporter.WriteModel(&system);
        const int steps = 4000;
        exporter.BeginResults(&system, steps);
        if (auto irr = std::dynamic_pointer_cast<IrregularWaves>(hydro->GetWave()))
            exporter.WriteIrregularInputs(irr->GetFrequenciesHz(), irr->GetSpectrum(), irr->GetFreeSurfaceTime(),
                                          irr->GetFreeSurfaceElevation());
        for (int i = 0; i < steps; ++i) {
            system.DoStepDynamics(timestep);
            exporter.RecordStep(&system);
        }
        exporter.SetRunMetadata("", "", 0.0, steps, timestep, system.GetChTime());
        exporter.Finalize();
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    std::cout << "Simulation finished." << std::endl;
    return 0;
}
