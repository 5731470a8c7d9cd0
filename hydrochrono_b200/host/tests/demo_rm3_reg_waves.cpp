// RM3 two-body point absorber in regular waves through the reference's API surface
// (cf. /root/reference/demos/rm3/demo_rm3_reg_waves.cpp:58-171): float + spar plate, prismatic joint between the two
// MOVING bodies, translational PTO damper, HHT stepper at dt = 0.01 s, RegularWave A = 1.0 m / omega = 2.10 rad/s,
// 12-DoF coupled radiation convolution.  rm3.h5 is stripped from the reference snapshot, so the tables are the
// synthetic RM3-shaped BEMIO file the pytest writes.
// usage: demo_rm3_reg_waves <rm3_like.h5> <out.txt> [duration = 40] [pto damping = 0] [results.h5]
// With HYDROC_STATE_TRACE=<file> TestHydro logs every evaluation (time, pose, velocity, gravity, total force).
#include <hydroc/hydro_forces.h>
#include <hydroc/simulation_exporter.h>

#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>

using namespace chrono;

int main(int argc, char* argv[]) {
    if (argc < 3) { std::cerr << "usage: demo_rm3_reg_waves <rm3.h5> <out.txt> [duration] [pto damping]" << std::endl; return 2; }
    const std::string h5fname = argv[1];
    const double simulationDuration = argc > 3 ? std::atof(argv[3]) : 40.0;
    const double pto_damping = argc > 4 ? std::atof(argv[4]) : 0.0;

    ChSystemNSC system;
    system.SetGravitationalAcceleration(ChVector3d(0.0, 0.0, -9.81));
    const double timestep = 0.01;
    system.SetTimestepperType(ChTimestepper::Type::HHT);
    system.SetSolverType(ChSolver::Type::GMRES);
    system.GetSolver()->AsIterative()->SetMaxIterations(300);

    std::shared_ptr<ChBody> float_body1 = chrono_types::make_shared<ChBodyEasyMesh>("float_cog.obj", 0, false, true, false);
    system.Add(float_body1);
    float_body1->SetName("body1");
    float_body1->SetPos(ChVector3d(0, 0, -0.72));
    float_body1->SetMass(725834);
    float_body1->SetInertiaXX(ChVector3d(20907301.0, 21306090.66, 37085481.11));

    std::shared_ptr<ChBody> plate_body2 = chrono_types::make_shared<ChBodyEasyMesh>("plate_cog.obj", 0, false, true, false);
    system.Add(plate_body2);
    plate_body2->SetName("body2");
    plate_body2->SetPos(ChVector3d(0, 0, -21.29));
    plate_body2->SetMass(886691);
    plate_body2->SetInertiaXX(ChVector3d(94419614.57, 94407091.24, 28542224.82));

    auto prismatic = chrono_types::make_shared<ChLinkLockPrismatic>();
    prismatic->Initialize(float_body1, plate_body2, false, ChFramed(ChVector3d(0, 0, -0.72)), ChFramed(ChVector3d(0, 0, -21.29)));
    system.AddLink(prismatic);

    auto prismatic_pto = chrono_types::make_shared<ChLinkTSDA>();
    prismatic_pto->Initialize(float_body1, plate_body2, false, ChVector3d(0, 0, -0.72), ChVector3d(0, 0, -21.29));
    prismatic_pto->SetDampingCoefficient(pto_damping);
    system.AddLink(prismatic_pto);

    std::vector<std::shared_ptr<ChBody>> bodies{float_body1, plate_body2};
    std::vector<double> time_vector, float_heave_position, float_drift_position, plate_heave_position;
    try {
        auto my_hydro_inputs = std::make_shared<RegularWave>(static_cast<unsigned int>(bodies.size()));
        my_hydro_inputs->regular_wave_amplitude_ = 1.0;
        my_hydro_inputs->regular_wave_omega_ = 2.10;
        TestHydro hydro_forces(bodies, h5fname);
        hydro_forces.AddWaves(my_hydro_inputs);

        std::unique_ptr<hydroc::SimulationExporter> exporter;
        if (argc > 5) {   // results file in the reference's schema v0.3, incl. the joint / TSDA channels
            hydroc::SimulationExporter::Options eo;
            eo.output_path = argv[5];
            eo.scenario_type = "regular"; eo.scenario_H = 2.0; eo.scenario_T = 2.0 * CH_PI / 2.10;
            exporter = std::make_unique<hydroc::SimulationExporter>(eo);
            exporter->WriteSimulationInfo(&system, CHRONO_VERSION, "rm3_reg_waves", timestep, simulationDuration);
            exporter->WriteModel(&system);
            exporter->BeginResults(&system, int(simulationDuration / timestep) + 2);
        }
        while (system.GetChTime() <= simulationDuration) {
            system.DoStepDynamics(timestep);
            if (exporter) exporter->RecordStep(&system);
            time_vector.push_back(system.GetChTime());
            float_heave_position.push_back(float_body1->GetPos().z());
            float_drift_position.push_back(float_body1->GetPos().x());
            plate_heave_position.push_back(plate_body2->GetPos().z());
        }
        if (exporter) exporter->Finalize();
        const HydroProfileStats st = hydro_forces.GetProfileStats();
        std::cout << "radiation_calls " << st.radiation_calls << " steps " << time_vector.size() << std::endl;
        // joint residuals: transverse offset and relative rotation of the float against the spar
        const ChVector3d d = plate_body2->GetRot().RotateBack(float_body1->GetPos() - plate_body2->GetPos());
        const ChQuaterniond q = plate_body2->GetRot().GetConjugate() * float_body1->GetRot();
        std::cout << std::setprecision(3) << std::scientific << "joint_transverse " << std::hypot(d.x(), d.y()) << " joint_rel_rot "
                  << std::sqrt(q.e1 * q.e1 + q.e2 * q.e2 + q.e3 * q.e3) << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }

    std::ofstream outputFile(argv[2]);
    if (!outputFile.is_open()) { std::cout << "Failed to open output file for writing" << std::endl; return 1; }
    outputFile << std::left << std::setw(10) << "Time (s)" << std::right << std::setw(16) << "Float Heave (m)" << std::right
               << std::setw(16) << "Plate Heave (m)" << std::right << std::setw(16) << "Float Drift (x) (m)" << std::endl;
    for (size_t i = 0; i < time_vector.size(); ++i)
        outputFile << std::left << std::setw(10) << std::setprecision(2) << std::fixed << time_vector[i] << std::right
                   << std::setw(16) << std::setprecision(6) << std::fixed << float_heave_position[i] << std::right << std::setw(16)
                   << std::setprecision(6) << std::fixed << plate_heave_position[i] << std::right << std::setw(16)
                   << std::setprecision(6) << std::fixed << float_drift_position[i] << std::endl;
    return 0;
}
