// API-surface test of the host layer on a two-body (RM3-shaped) system: the calls a HydroChrono user makes, with
// the semantics the reference documents.  Needs a CUDA device.
// usage: test_api_surface <rm3_like.h5> <hydro.yaml>
#include <hydroc/chloadaddedmass.h>
#include <hydroc/helper.h>
#include <hydroc/hydro_forces.h>
#include <hydroc/hydro_yaml_parser.h>
#include <hydroc/setup_hydro_from_yaml.h>

#include <cmath>
#include <iostream>

static int failures = 0;
#define CHECK(cond)                                                                                            \
    do {                                                                                                       \
        if (!(cond)) { std::cerr << "CHECK failed: " #cond " at line " << __LINE__ << std::endl; ++failures; } \
    } while (0)

using namespace chrono;

static std::shared_ptr<ChBody> make_body(ChSystem& sys, const std::string& name, double mass, double z) {
    auto b = chrono_types::make_shared<ChBody>();
    b->SetName(name);
    b->SetMass(mass);
    b->SetInertiaXX(ChVector3d(2.0e7, 2.1e7, 3.7e7));
    b->SetPos(ChVector3d(0, 0, z));
    sys.Add(b);
    return b;
}

int main(int argc, char* argv[]) {
    if (argc < 3) { std::cerr << "usage: test_api_surface <rm3_like.h5> <hydro.yaml>" << std::endl; return 2; }
    const std::string h5 = argv[1];
    try {
        // ---- H5FileInfo / HydroData getters ----
        HydroData hd = H5FileInfo(h5, 2).ReadH5Data();
        CHECK(hd.GetRIRFDims(0) == 6 && hd.GetRIRFDims(1) == 12 && hd.GetRIRFDims(2) > 10);
        CHECK(hd.GetRhoVal() == 1000.0);
        CHECK(hd.GetBodyInfos().size() == 2 && hd.GetBodyInfos()[1].body_name == "body2");
        CHECK(hd.GetInfAddedMassMatrix(1).rows() == 6 && hd.GetInfAddedMassMatrix(1).cols() == 12);
        CHECK(hd.GetRIRFTimeVector().size() == hd.GetRIRFDims(2));
        CHECK(hd.GetHydrostaticStiffnessVal(0, 2, 2) == hd.GetLinMatrix(0)(2, 2) * 1000.0 * 9.81);
        CHECK(std::abs(hd.GetCBVector(0)[2] - hd.GetCGVector(0)[2] - 0.1) < 1e-12);

        // ---- a two-body system with an extra, non-hydro body added AFTER the hydro bodies ----
        ChSystemNSC system;
        system.SetGravitationalAcceleration(ChVector3d(0.0, 0.0, -9.81));
        auto b1 = make_body(system, "body1", 725834.0, -0.72 + 0.1);
        auto b2 = make_body(system, "body2", 886691.0, -21.29 + 0.1);
        auto extra = make_body(system, "payload", 100.0, 5.0);
        std::vector<std::shared_ptr<ChBody>> bodies{b1, b2};

        auto reg = std::make_shared<RegularWave>(2);
        reg->regular_wave_amplitude_ = 1.0;     // demos/rm3/demo_rm3_reg_waves.cpp: A = 1.0 m, omega = 2.10 rad/s
        reg->regular_wave_omega_ = 2.10;
        TestHydro hydro(bodies, h5, reg);

        // first evaluation at t = 0 through the Chrono hook chain: ComponentFunc -> ForceFunc6d -> TestHydro
        const double f_heave_1 = hydro.CoordinateFuncForBody(1, 2);
        const double f_heave_2 = hydro.CoordinateFuncForBody(2, 2);
        CHECK(std::isfinite(f_heave_1) && std::isfinite(f_heave_2));
        std::vector<double> hs = hydro.ComputeForceHydrostatics();
        std::vector<double> rad = hydro.ComputeForceRadiationDampingConv();
        Eigen::VectorXd wv = hydro.ComputeForceWaves();
        CHECK(hs.size() == 12 && rad.size() == 12 && wv.size() == 12);
        CHECK(f_heave_1 == hs[2] - rad[2] + wv[2]);                 // total = hs - rad + waves
        CHECK(f_heave_2 == hs[8] - rad[8] + wv[8]);
        for (double r : rad) CHECK(r == 0.0);                       // a single history entry: no convolution yet
        // WaveBase::GetForceAtTime agrees with the wave component of the step
        Eigen::VectorXd wf = hydro.GetWave()->GetForceAtTime(0.0);
        for (int i = 0; i < 12; ++i) CHECK(wf[i] == wv[i]);
        // RegularWave phase quirk: body 2 uses body 1's interpolated phases
        Eigen::VectorXd ph = reg->GetExcitationPhase(), mg = reg->GetExcitationMag();
        CHECK(std::abs(wv[8] - mg[8] * 1.0 * std::cos(2.10 * 0.0 + ph[2])) <= 1e-9 * std::abs(wv[8]) + 1e-12);
        // index errors keep their exception types
        bool threw = false;
        try { hydro.CoordinateFuncForBody(3, 0); } catch (const std::out_of_range&) { threw = true; }
        CHECK(threw);
        threw = false;
        try { hydro.GetRIRFval(12, 0, 0); } catch (const std::out_of_range&) { threw = true; }
        CHECK(threw);

        // ---- added-mass load: Jacobian padded to the system size, block at (0,0); R += c M w ----
        CHECK(system.GetNumCoordsVelLevel() == 18);
        system.DoStepDynamics(0.01);                                // assembles the load Jacobian
        HydroProfileStats st = hydro.GetProfileStats();
        CHECK(st.radiation_calls == 1 && st.hydrostatics_calls == 1 && st.waves_calls == 1);
        for (int i = 0; i < 20; ++i) system.DoStepDynamics(0.01);
        CHECK(std::abs(system.GetChTime() - 0.21) < 1e-12);
        rad = hydro.ComputeForceRadiationDampingConv();             // evaluates at the current time (0.21)
        double radnorm = 0;
        for (double r : rad) radnorm += r * r;
        CHECK(radnorm > 0.0);
        CHECK(extra->GetPos().z() < 5.0);                           // the non-hydro body just falls
        CHECK(hydro.GetProfileStats().radiation_calls == 22);

        // ---- TaperedDirect re-stages the kernel ----
        const double k_before = hydro.GetRIRFval(2, 2, hd.GetRIRFDims(2) - 1);
        hydro.SetRadiationConvolutionMode(TestHydro::RadiationConvolutionMode::TaperedDirect);
        TestHydro::TaperedDirectOptions opts;
        hydro.SetTaperedDirectOptions(opts);
        CHECK(k_before != 0.0 && std::abs(hydro.GetRIRFval(2, 2, hd.GetRIRFDims(2) - 1)) < 1e-2 * std::abs(k_before));
        CHECK(hydro.GetRIRFval(2, 2, 0) == hd.GetRIRFVal(0, 2, 2, 0));   // first samples are copied as they are

        // ---- YAML front end ----
        ChSystemNSC sys2;
        sys2.SetGravitationalAcceleration(ChVector3d(0.0, 0.0, -9.8));
        auto c1 = make_body(sys2, "body1", 725834.0, -0.72);
        auto c2 = make_body(sys2, "body2", 886691.0, -21.29);
        YAMLHydroData y = ReadHydroYAML(argv[2]);
        CHECK(y.bodies.size() == 2 && y.waves.type == "irregular");
        std::unique_ptr<TestHydro> h2 = SetupHydroFromYAML(y, {c1, c2}, 0.05, 4.0, 1.0);
        for (int i = 0; i < 10; ++i) sys2.DoStepDynamics(0.05);
        auto irr = std::static_pointer_cast<IrregularWaves>(h2->GetWave());
        CHECK(irr->GetParams().seed_ == 5 && irr->GetParams().peak_enhancement_factor_ == 1.0);
        CHECK(irr->GetSpectrum().size() == irr->GetFrequenciesHz().size() && !irr->GetSpectrum().empty());
        CHECK(irr->GetFreeSurfaceElevation().size() == irr->GetFreeSurfaceTime().size());
        const double eta0 = irr->GetElevation(Eigen::Vector3d(0, 0, 0), 0.5);
        CHECK(std::isfinite(eta0));
        Eigen::Vector3d vel = irr->GetVelocity(Eigen::Vector3d(0, 0, -1.0), 0.5);
        CHECK(std::isfinite(vel.x()) && vel.y() == 0.0);
        Eigen::VectorXd we = h2->ComputeForceWaves();
        double wn = 0;
        for (int i = 0; i < 12; ++i) wn += we[i] * we[i];
        CHECK(wn > 0.0);
        threw = false;
        try { SetupHydroFromYAML(y, {extra}, 0.05, 4.0, 1.0); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);                                               // "No hydrodynamic bodies found in Chrono system"

        // get_lower_index contract (src/helper.cpp:8-22)
        std::vector<double> ticks{0, 1, 2, 3, 4};
        CHECK(get_lower_index(2.5, ticks) == 2 && get_lower_index(2.0, ticks) == 1);
        threw = false;
        try { get_lower_index(0.5, ticks); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    if (failures) { std::cerr << failures << " check(s) failed" << std::endl; return 1; }
    std::cout << "API surface test passed" << std::endl;
    return 0;
}
