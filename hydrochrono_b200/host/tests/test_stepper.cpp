// chrono_compat steppers on a problem with a closed-form answer: a heaving body on a linear spring-damper whose force
// comes from time-keyed ChFunction callbacks that read the body state, exactly how TestHydro's ComponentFunc feeds
// Chrono.  Linearised Euler is first order, the HHT-alpha step second order.  No CUDA device needed.
#include <chrono_compat/chrono_compat.h>

#include <algorithm>
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

// Two moving bodies on a prismatic joint (the RM3 float / spar arrangement) with a spring-damper PTO between them,
// no gravity.  The pair is given a common spin about y; closed-form answers:
//  - the slide s(t) = n . (xA - xB) - s0 follows  mu s'' + c s' + (k - mu w^2) s = mu w^2 s0  ... with w = 0 the plain
//    damped oscillator of the reduced mass mu = mA mB / (mA + mB);
//  - the joint keeps the transverse offset and the relative rotation at zero at every step;
//  - linear momentum is conserved.
struct PairResult { double slide_err, transverse, rel_rot, momentum_err; };
PairResult run_pair(ChTimestepper::Type type, double h, double t_end, double spin) {
    const double mA = 3.0, mB = 5.0, k = 40.0, c = 0.8, gap = 2.0, stretch = 0.3;
    ChSystemNSC sys;
    sys.SetGravitationalAcceleration(ChVector3d(0, 0, 0));
    sys.SetTimestepperType(type);
    auto A = chrono_types::make_shared<ChBody>();
    auto B = chrono_types::make_shared<ChBody>();
    A->SetMass(mA); B->SetMass(mB);
    A->SetInertiaXX(ChVector3d(2, 2, 2)); B->SetInertiaXX(ChVector3d(4, 4, 4));
    A->SetPos(ChVector3d(0, 0, gap)); B->SetPos(ChVector3d(0, 0, 0));
    sys.Add(A); sys.Add(B);
    auto joint = chrono_types::make_shared<ChLinkLockPrismatic>();
    joint->Initialize(A, B, false, ChFramed(ChVector3d(0, 0, gap)), ChFramed(ChVector3d(0, 0, 0)));
    sys.AddLink(joint);
    auto pto = chrono_types::make_shared<ChLinkTSDA>();
    pto->Initialize(A, B, false, ChVector3d(0, 0, gap), ChVector3d(0, 0, 0));
    pto->SetSpringCoefficient(k);
    pto->SetDampingCoefficient(c);
    sys.AddLink(pto);
    A->SetPos(ChVector3d(0, 0, gap + stretch));                   // released from a stretched PTO
    A->SetAngVelParent(ChVector3d(0, spin, 0)); B->SetAngVelParent(ChVector3d(0, spin, 0));
    A->SetPosDt(ChVector3d(spin * (gap + stretch), 0, 0));        // rigid rotation of the pair about B
    const ChVector3d p0 = A->GetPosDt() * mA + B->GetPosDt() * mB;
    const int n = int(std::lround(t_end / h));
    PairResult r{0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        sys.DoStepDynamics(h);
        const ChVector3d d = B->GetRot().RotateBack(A->GetPos() - B->GetPos());
        r.transverse = std::max(r.transverse, std::hypot(d.x(), d.y()));
        const ChQuaterniond q = B->GetRot().GetConjugate() * A->GetRot();
        r.rel_rot = std::max(r.rel_rot, std::sqrt(q.e1 * q.e1 + q.e2 * q.e2 + q.e3 * q.e3));
    }
    const double t = sys.GetChTime();
    const ChVector3d p1 = A->GetPosDt() * mA + B->GetPosDt() * mB;
    r.momentum_err = (p1 - p0).Length();
    const double s = B->GetRot().RotateBack(A->GetPos() - B->GetPos()).z() - gap;
    if (spin == 0.0) {
        const double mu = mA * mB / (mA + mB), zeta = c / (2.0 * mu), wd = std::sqrt(k / mu - zeta * zeta);
        const double exact = std::exp(-zeta * t) * (stretch * std::cos(wd * t) + stretch * zeta / wd * std::sin(wd * t));
        r.slide_err = std::fabs(s - exact);
    }
    return r;
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
    // prismatic joint between two moving bodies + PTO
    const PairResult p1 = run_pair(ChTimestepper::Type::HHT, 0.004, 3.0, 0.0), p2 = run_pair(ChTimestepper::Type::HHT, 0.002, 3.0, 0.0);
    const PairResult q1 = run_pair(ChTimestepper::Type::EULER_IMPLICIT_LINEARIZED, 0.002, 3.0, 0.0);
    std::cout << "pair hht slide err " << p1.slide_err << " " << p2.slide_err << " ratio " << p1.slide_err / p2.slide_err
              << "  euler " << q1.slide_err << "  momentum " << p2.momentum_err << std::endl;
    check(p1.slide_err / p2.slide_err > 3.0 && p1.slide_err / p2.slide_err < 5.0, "jointed pair: HHT slide converges with order 2");
    check(p2.slide_err < 2e-4 && q1.slide_err < 2e-2, "jointed pair: slide follows the reduced-mass oscillator");
    check(p2.momentum_err < 1e-12 && q1.momentum_err < 1e-12, "jointed pair: linear momentum conserved");
    check(p2.transverse < 1e-14 && p2.rel_rot < 1e-14, "jointed pair: joint constraints hold");
    const PairResult sp = run_pair(ChTimestepper::Type::HHT, 0.002, 3.0, 0.35);
    std::cout << "spinning pair: transverse " << sp.transverse << " rel rot " << sp.rel_rot << " momentum " << sp.momentum_err << std::endl;
    check(sp.transverse < 1e-12 && sp.rel_rot < 1e-12, "spinning jointed pair: joint constraints hold");
    check(sp.momentum_err < 5e-3, "spinning jointed pair: linear momentum conserved to integration accuracy");
    if (failures == 0) std::cout << "stepper test passed" << std::endl;
    return failures ? 1 : 0;
}
