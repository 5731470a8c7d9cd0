// chrono_compat steppers on a problem with a closed-form answer: a heaving body on a linear spring-damper whose force
// comes from time-keyed ChFunction callbacks that read the body state, exactly how TestHydro's ComponentFunc feeds
// Chrono.  Linearised Euler is first order, the HHT-alpha step second order.  No CUDA device needed.
#include <chrono_compat/chrono_compat.h>

#include <cmath>
#include <iostream>

using namespace chrono;

namespace {
struct SpringFn : ChFunction {
    ChBody* body;
    double k, c;
    SpringFn(ChBody* b, double k_, double c_) : body(b), k(k_), c(c_) {}
    ChFunction* Clone() const override { return new SpringFn(*this); }
    double GetVal(double) const override { return -k * body->GetPos().z() - c * body->GetPosDt().z(); }
};

double run(ChTimestepper::Type type, double h, double t_end) {
    const double m = 2.0, k = 50.0, c = 0.4, z0 = 1.0;
    ChSystemNSC sys;
    sys.SetGravitationalAcceleration(ChVector3d(0, 0, 0));
    sys.SetTimestepperType(type);
    auto b = chrono_types::make_shared<ChBody>();
    b->SetMass(m);
    b->SetPos(ChVector3d(0, 0, z0));
    sys.Add(b);
    auto f = chrono_types::make_shared<ChForce>();
    f->SetMode(ChForce::ForceType::FORCE);
    f->SetF_z(chrono_types::make_shared<SpringFn>(b.get(), k, c));
    b->AddForce(f);
    const int n = int(std::lround(t_end / h));
    for (int i = 0; i < n; ++i) sys.DoStepDynamics(h);
    const double t = sys.GetChTime();
    const double zeta = c / (2.0 * m), wd = std::sqrt(k / m - zeta * zeta);
    const double exact = std::exp(-zeta * t) * (z0 * std::cos(wd * t) + z0 * zeta / wd * std::sin(wd * t));
    return std::fabs(b->GetPos().z() - exact);
}
}  // namespace

int main() {
    int failures = 0;
    auto check = [&](bool ok, const char* what) {
        if (!ok) { std::cerr << "CHECK failed: " << what << std::endl; ++failures; }
    };
    const double T = 4.0;
    const double e1 = run(ChTimestepper::Type::EULER_IMPLICIT_LINEARIZED, 0.004, T);
    const double e2 = run(ChTimestepper::Type::EULER_IMPLICIT_LINEARIZED, 0.002, T);
    const double h1 = run(ChTimestepper::Type::HHT, 0.004, T);
    const double h2 = run(ChTimestepper::Type::HHT, 0.002, T);
    std::cout << "euler " << e1 << " " << e2 << " ratio " << e1 / e2 << "\n"
              << "hht   " << h1 << " " << h2 << " ratio " << h1 / h2 << std::endl;
    check(e1 / e2 > 1.6 && e1 / e2 < 2.6, "linearised Euler converges with order 1");
    check(h1 / h2 > 3.0 && h1 / h2 < 5.0, "HHT-alpha step converges with order 2");
    check(h1 < 0.2 * e1, "HHT more accurate than Euler at the same step");
    if (failures == 0) std::cout << "stepper test passed" << std::endl;
    return failures ? 1 : 0;
}
