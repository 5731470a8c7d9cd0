// hydro.yaml period sweep -> ensemble instances (/root/reference/src/hydro_yaml_parser.cpp:441-524 parses
// waves.period.{values,linspace,range} into WaveSettings::period_values, which the reference's SetupHydroFromYAML never
// consumes).  One sphere system per sweep point (x seed), all advanced in lock-step around ONE batched device
// evaluation per time value; the same cases are then run one by one through the single-system TestHydro and the
// trajectories compared.
// usage: demo_sweep_yaml <hydro.yaml> <out.txt> <euler|hht> [duration = 30] [seeds per period = 1]
#include <hydroc/chloadaddedmass.h>
#include <hydroc/hydro_ensemble.h>
#include <hydroc/hydro_forces.h>
#include <hydroc/hydro_yaml_parser.h>
#include <hydroc/setup_hydro_from_yaml.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>

using namespace chrono;

struct Sphere {
    ChSystemNSC sys;
    std::shared_ptr<ChBody> ground, body;
    explicit Sphere(bool hht) {
        sys.SetGravitationalAcceleration(ChVector3d(0.0, 0.0, -9.81));
        if (hht) sys.SetTimestepperType(ChTimestepper::Type::HHT);
        ground = chrono_types::make_shared<ChBody>();
        sys.AddBody(ground);
        ground->SetPos(ChVector3d(0, 0, -5));
        ground->SetFixed(true);
        body = chrono_types::make_shared<ChBodyEasyMesh>("oes_task10_sphere.obj", 1000, false, true, false);
        sys.Add(body);
        body->SetName("body1");
        body->SetPos(ChVector3d(0, 0, -2));
        body->SetMass(261.8e3);
        auto prismatic = chrono_types::make_shared<ChLinkLockPrismatic>();
        prismatic->Initialize(body, ground, false, ChFramed(ChVector3d(0, 0, -2)), ChFramed(ChVector3d(0, 0, -5)));
        sys.AddLink(prismatic);
    }
};

int main(int argc, char* argv[]) {
    if (argc < 4) { std::cerr << "usage: demo_sweep_yaml <hydro.yaml> <out.txt> <euler|hht> [duration] [seeds]" << std::endl; return 2; }
    const bool hht = std::strcmp(argv[3], "hht") == 0;
    const double duration = argc > 4 ? std::atof(argv[4]) : 30.0;
    const int seeds_per_period = argc > 5 ? std::atoi(argv[5]) : 1;
    const double dt = 0.015, ramp = 3.0;
    try {
        YAMLHydroData y = ReadHydroYAML(argv[1]);
        const int P = std::max<size_t>(1, y.waves.period_values.size());
        const int B = P * seeds_per_period;
        const int nsteps = int(std::lround(duration / dt));

        // ---- the sweep as ONE ensemble ----
        std::vector<std::unique_ptr<Sphere>> inst;
        std::vector<TestHydroEnsemble::BodyList> lists;
        std::vector<ChSystem*> systems;
        for (int i = 0; i < B; ++i) {
            inst.push_back(std::make_unique<Sphere>(hht));
            lists.push_back({inst.back()->body});
            systems.push_back(&inst.back()->sys);
        }
        std::unique_ptr<TestHydroEnsemble> ens = SetupHydroSweepFromYAML(y, lists, dt, duration, ramp, seeds_per_period);
        std::vector<std::vector<double>> z(B, std::vector<double>(nsteps));
        for (int n = 0; n < nsteps; ++n) {
            ChSystem::DoStepDynamicsLockstep(systems, dt);
            for (int i = 0; i < B; ++i) z[i][n] = inst[i]->body->GetPos().z();
        }
        std::cout << "instances " << B << " steps " << nsteps << " device_evaluations " << ens->DeviceEvaluations() << std::endl;
        {   // batched added-mass residual on the device against each system's own ChLoadAddedMass (host) load
            const int nsys = 6;
            std::vector<double> w(size_t(B) * nsys), R(size_t(B) * nsys, 0.5);
            for (size_t k = 0; k < w.size(); ++k) w[k] = 0.01 * double(k % 17) - 0.05;
            ens->AddedMassMvAll(nsys, 1.75, w, R);
            double worst_mv = 0.0;
            std::vector<std::shared_ptr<ChLoadable>> loadables{inst[0]->body};
            ChLoadAddedMass ref(ens->GetHydroData().GetBodyInfos(), loadables, &inst[0]->sys);
            ref.ComputeJacobian(nullptr, nullptr);
            for (int i = 0; i < B; ++i) {
                ChVectorDynamic<> Rv(nsys), wv(nsys);
                for (int k = 0; k < nsys; ++k) { Rv(k) = 0.5; wv(k) = w[size_t(i) * nsys + k]; }
                ref.LoadIntLoadResidual_Mv(Rv, wv, 1.75);
                for (int k = 0; k < nsys; ++k) worst_mv = std::max(worst_mv, std::fabs(Rv(k) - R[size_t(i) * nsys + k]) / (1.0 + std::fabs(Rv(k))));
            }
            std::cout << std::scientific << "added_mass_mv_batched_rel_diff " << worst_mv << std::endl;
        }

        // ---- the same cases one at a time through the single-system TestHydro ----
        double worst = 0.0;
        for (int i = 0; i < B; ++i) {
            Sphere s(hht);
            YAMLHydroData yi = y;
            yi.waves.period = y.waves.period_values.empty() ? y.waves.period : y.waves.period_values[i % P];
            yi.waves.seed = (y.waves.seed > 0 ? y.waves.seed : 1) + i / P;
            std::unique_ptr<TestHydro> h = SetupHydroFromYAML(yi, {s.body}, dt, duration, ramp);
            for (int n = 0; n < nsteps; ++n) {
                s.sys.DoStepDynamics(dt);
                worst = std::max(worst, std::fabs(s.body->GetPos().z() - z[i][n]));
            }
        }
        std::cout << std::scientific << std::setprecision(3) << "max_abs_diff_vs_single_runs " << worst << std::endl;

        std::ofstream out(argv[2]);
        out << std::setprecision(17);
        for (int n = 0; n < nsteps; ++n) {
            out << (n + 1) * dt;
            for (int i = 0; i < B; ++i) out << ' ' << z[i][n];
            out << '\n';
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
