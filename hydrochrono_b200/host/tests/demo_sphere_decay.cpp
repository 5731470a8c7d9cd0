// Sphere heave decay through the reference's own API surface (cf. tests/regression/sphere/demo_sphere_decay.cpp):
// ChSystemNSC, one ChBody "body1", TestHydro(bodies, h5), NoWave, DoStepDynamics in a loop, heave written per step.
// usage: demo_sphere_decay <sphere.h5> <out.txt>
#include <hydroc/helper.h>
#include <hydroc/hydro_forces.h>

#include <iostream>

#include "common_demo.h"

using namespace chrono;

int main(int argc, char* argv[]) {
    if (argc < 3) { std::cerr << "usage: demo_sphere_decay <sphere.h5> <out.txt>" << std::endl; return 2; }
    const std::string h5fname = argv[1];

    ChSystemNSC system;
    system.SetGravitationalAcceleration(ChVector3d(0.0, 0.0, -9.81));
    double timestep = 0.015;
    system.SetSolverType(ChSolver::Type::GMRES);
    system.GetSolver()->AsIterative()->SetMaxIterations(300);
    double simulationDuration = 40.0;

    std::vector<double> time_vector, heave_position;

    std::shared_ptr<ChBody> sphereBody = chrono_types::make_shared<ChBodyEasyMesh>("oes_task10_sphere.obj", 1000, false, true, false);
    sphereBody->SetName("body1");   // must match the .h5 body name
    sphereBody->SetPos(ChVector3d(0, 0, -1));
    sphereBody->SetMass(261.8e3);
    system.Add(sphereBody);

    auto default_dont_add_waves = std::make_shared<NoWave>(1);
    std::vector<std::shared_ptr<ChBody>> bodies;
    bodies.push_back(sphereBody);

    try {
        TestHydro hydro_forces(bodies, h5fname);
        hydro_forces.AddWaves(default_dont_add_waves);

        while (system.GetChTime() <= simulationDuration) {
            system.DoStepDynamics(timestep);
            time_vector.push_back(system.GetChTime());
            heave_position.push_back(sphereBody->GetPos().z());
        }
        HydroProfileStats st = hydro_forces.GetProfileStats();
        std::cout << "steps " << time_vector.size() << " radiation_calls " << st.radiation_calls << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    if (write_heave(argv[2], time_vector, heave_position)) return 1;
    std::cout << "Simulation finished." << std::endl;
    return 0;
}
