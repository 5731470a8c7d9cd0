// Water kinematics through the reference's class surface: RegularWave / IrregularWaves attached to a
// TestHydro, GetElevation / GetVelocity / GetAcceleration at a list of points and times, written as text for the
// pytest that holds them against the oracle's restatement of src/wave_types.cpp:14-160,301-313,515-550.
// usage: test_kinematics <sphere.h5> <out.txt>
#include <hydroc/hydro_forces.h>

#include <fstream>
#include <iomanip>
#include <iostream>

using namespace chrono;

static void dump(std::ofstream& out, const char* tag, WaveBase& w) {
    const double pts[5][3] = {{0.0, 0.0, 0.0}, {3.5, 1.0, -2.0}, {-12.0, 0.0, -7.5}, {40.0, -3.0, -0.25}, {1.0, 0.0, 0.4}};
    const double times[4] = {0.0, 0.37, 11.2, 95.0};
    for (auto& p : pts)
        for (double t : times) {
            const Eigen::Vector3d pos(p[0], p[1], p[2]);
            const double eta = w.GetElevation(pos, t);
            const Eigen::Vector3d v = w.GetVelocity(pos, t), a = w.GetAcceleration(pos, t);
            out << tag << ' ' << std::setprecision(17) << p[0] << ' ' << p[1] << ' ' << p[2] << ' ' << t << ' ' << eta << ' '
                << v[0] << ' ' << v[1] << ' ' << v[2] << ' ' << a[0] << ' ' << a[1] << ' ' << a[2] << '\n';
        }
}

int main(int argc, char* argv[]) {
    if (argc < 3) { std::cerr << "usage: test_kinematics <sphere.h5> <out.txt>" << std::endl; return 2; }
    try {
        std::ofstream out(argv[2]);
        auto make = [&](ChSystemNSC& sys) {
            sys.SetGravitationalAcceleration(ChVector3d(0.0, 0.0, -9.81));
            auto body = chrono_types::make_shared<ChBodyEasyMesh>("sphere.obj", 1000, false, true, false);
            sys.Add(body);
            body->SetName("body1");
            body->SetPos(ChVector3d(0, 0, -2));
            body->SetMass(261.8e3);
            return std::vector<std::shared_ptr<ChBody>>{body};
        };
        {   // regular wave (src/wave_types.cpp:301-313)
            ChSystemNSC sys;
            auto bodies = make(sys);
            auto w = std::make_shared<RegularWave>(1u);
            w->regular_wave_amplitude_ = 0.8; w->regular_wave_omega_ = 0.9; w->regular_wave_phase_ = 0.4;
            TestHydro hydro(bodies, argv[1]);
            hydro.AddWaves(w);
            dump(out, "regular", *w);
            w->mwl_ = 0.3;
            dump(out, "regular_mwl", *w);
        }
        for (int stretch = 0; stretch < 2; ++stretch) {   // irregular waves (:515-550), Wheeler stretching off / on
            ChSystemNSC sys;
            auto bodies = make(sys);
            IrregularWaveParams p;
            p.num_bodies_ = 1; p.simulation_dt_ = 0.05; p.simulation_duration_ = 20.0; p.ramp_duration_ = 0.0;
            p.wave_height_ = 2.0; p.wave_period_ = 9.0; p.frequency_min_ = 0.02; p.frequency_max_ = 0.6; p.nfrequencies_ = 60;
            p.peak_enhancement_factor_ = 3.3; p.seed_ = 7; p.wave_stretching_ = stretch != 0;
            auto w = std::make_shared<IrregularWaves>(p);
            TestHydro hydro(bodies, argv[1]);
            hydro.AddWaves(w);
            dump(out, stretch ? "irregular_wheeler" : "irregular", *w);
            w->mwl_ = -0.2;
            dump(out, stretch ? "irregular_wheeler_mwl" : "irregular_mwl", *w);
        }
        std::cout << "kinematics written" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
