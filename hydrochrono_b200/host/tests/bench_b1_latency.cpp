// Latency of ONE force evaluation for one system (B = 1), irregular waves, history window full -- the drop-in case:
//  (1) hc_step through the C ABI from C++ (host buffers in, forces out, synchronous);
//  (2) the same evaluation through the reference's class surface: the first ComponentFunc::GetVal at a new ChTime
//      (state gather + hc_step) and the five cache hits that follow it (src/hydro_forces.cpp:742-767).
// usage: bench_b1_latency <tables.h5> <num_bodies> <dt> <prefill_steps> <timed_steps>
#include <hydroc/hydro_forces.h>
#include <hydrochrono_b200.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

using namespace chrono;
using clk = std::chrono::steady_clock;

static double median(std::vector<double> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }
static double pct(std::vector<double> v, double p) { std::sort(v.begin(), v.end()); return v[size_t(p * (v.size() - 1))]; }

int main(int argc, char* argv[]) {
    if (argc < 6) { std::cerr << "usage: bench_b1_latency <tables.h5> <num_bodies> <dt> <prefill> <timed>" << std::endl; return 2; }
    const int N = std::atoi(argv[2]), D = 6 * N;
    const double dt = std::atof(argv[3]);
    const int prefill = std::atoi(argv[4]), timed = std::atoi(argv[5]);
    try {
        // ---- (1) C ABI ----
        hc_tables* T = nullptr;
        if (hc_tables_load_h5(argv[1], N, &T) != HC_OK) throw std::runtime_error(hc_last_error());
        hc_ensemble_opts o;
        hc_ensemble_default_opts(&o);
        o.batch = 1; o.dt_hint = dt;
        hc_ensemble* E = nullptr;
        if (hc_ensemble_create(T, &o, &E) != HC_OK) throw std::runtime_error(hc_last_error());
        hc_irregular_params q;
        hc_irregular_default_params(&q);
        q.simulation_dt = dt; q.simulation_duration = (prefill + timed + 64) * dt; q.ramp_duration = 20.0;
        q.wave_height = 2.5; q.wave_period = 8.0; q.peak_enhancement_factor = 3.3; q.nfrequencies = 1000; q.seed = 1;
        if (hc_waves_irregular(E, &q, nullptr, nullptr, nullptr) != HC_OK) throw std::runtime_error(hc_last_error());
        std::vector<double> pose(D), vel(D), force(D), lat;
        const double g[3] = {0.0, 0.0, -9.81};
        double t = 0.0, checksum = 0.0;
        for (int n = 0; n < prefill + timed; ++n) {
            for (int d = 0; d < D; ++d) { pose[d] = 0.05 * std::sin(0.7 * t + d); vel[d] = 0.035 * std::cos(0.7 * t + d); }
            const auto t0 = clk::now();
            if (hc_step(E, t, pose.data(), vel.data(), g, force.data(), nullptr) != HC_OK) throw std::runtime_error(hc_last_error());
            if (n >= prefill) lat.push_back(std::chrono::duration<double, std::micro>(clk::now() - t0).count());
            checksum += force[2];
            t += dt;
        }
        std::printf("C ABI hc_step, B = 1, D = %d: median %.1f us  p10 %.1f  p90 %.1f  (checksum %.6e)\n", D, median(lat),
                    pct(lat, 0.1), pct(lat, 0.9), checksum);
        hc_ensemble_destroy(E);
        hc_tables_destroy(T);

        // ---- (2) class surface: TestHydro + ForceFunc6d, time advanced by hand (no integrator in the timing) ----
        ChSystemNSC system;
        system.SetGravitationalAcceleration(ChVector3d(0.0, 0.0, -9.81));
        std::vector<std::shared_ptr<ChBody>> bodies;
        for (int b = 0; b < N; ++b) {
            auto body = chrono_types::make_shared<ChBody>();
            body->SetName("body" + std::to_string(b + 1));
            body->SetMass(5e5);
            system.AddBody(body);
            bodies.push_back(body);
        }
        IrregularWaveParams p;
        p.num_bodies_ = N; p.simulation_dt_ = dt; p.simulation_duration_ = q.simulation_duration; p.ramp_duration_ = 20.0;
        p.wave_height_ = 2.5; p.wave_period_ = 8.0; p.peak_enhancement_factor_ = 3.3; p.nfrequencies_ = 1000; p.seed_ = 1;
        auto waves = std::make_shared<IrregularWaves>(p);
        TestHydro hydro(bodies, argv[1]);
        hydro.AddWaves(waves);
        std::vector<double> first, rest;
        t = 0.0;
        for (int n = 0; n < prefill + timed; ++n) {
            system.SetChTime(t);
            for (int b = 0; b < N; ++b) {
                bodies[b]->SetPos(ChVector3d(0.05 * std::sin(0.7 * t + b), 0.0, 0.05 * std::cos(0.7 * t + b)));
                bodies[b]->SetPosDt(ChVector3d(0.035 * std::cos(0.7 * t + b), 0.0, -0.035 * std::sin(0.7 * t + b)));
            }
            const auto t0 = clk::now();
            const double f0 = hydro.CoordinateFuncForBody(1, 2);            // first evaluation at this time: one hc_step
            const auto t1 = clk::now();
            double s = f0;
            for (int b = 1; b <= N; ++b)
                for (int k = 0; k < 6; ++k) s += hydro.CoordinateFuncForBody(b, k);   // cache hits
            const auto t2 = clk::now();
            if (n >= prefill) {
                first.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
                rest.push_back(std::chrono::duration<double, std::micro>(t2 - t1).count());
            }
            checksum += s;
            t += dt;
        }
        std::printf("TestHydro::CoordinateFuncForBody, first call at a new time: median %.1f us  p90 %.1f;  the %d cached calls "
                    "that follow: %.2f us in total  (checksum %.6e)\n", median(first), pct(first, 0.9), 6 * N, median(rest), checksum);
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
