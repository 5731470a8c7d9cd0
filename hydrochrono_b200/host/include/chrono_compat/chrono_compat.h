// chrono_compat -- the slice of Project Chrono's API that HydroChrono's hydro plugin touches, plus a
// small rigid-body stepper, so that the host layer (hydroc/*.h) and the reference-style demo mains can be built
// and run where Project Chrono is not installed.  It is NOT a general multibody engine: bodies are 6-DoF rigid bodies,
// a ChLinkLockPrismatic either locks a body to translation along the joint axis relative to a fixed body or makes it
// the 1-DoF child of another moving body (reduced coordinates: the pair moves as one rigid body plus the relative
// slide -- the RM3 float / spar arrangement), a ChLinkTSDA adds a spring-damper (the PTO);
// ChSystem::DoStepDynamics advances with the linearised-Euler scheme the reference's sphere goldens were produced
// with (force evaluated once at (t_n, x_n, v_n), v += dt (M + M_added)^-1 F, x += dt v), or, after
// SetTimestepperType(HHT) as in the reference's YAML runs, with an HHT-alpha step (chrono_compat.cpp).
//
// With the real Chrono, compile the host layer with -DHYDROC_HAVE_CHRONO and this header is not used.
#pragma once
#ifdef HYDROC_HAVE_CHRONO
#include <chrono/physics/ChBody.h>
#include <chrono/physics/ChForce.h>
#include <chrono/physics/ChLoad.h>
#include <chrono/physics/ChLoadContainer.h>
#include <chrono/physics/ChSystemNSC.h>
#else
#include <cmath>
#include <memory>
#include <string>
#include <vector>

namespace chrono_types {
template <class T, class... Args>
std::shared_ptr<T> make_shared(Args&&... args) { return std::make_shared<T>(std::forward<Args>(args)...); }
}  // namespace chrono_types

namespace chrono {

constexpr double CH_PI = 3.14159265358979323846;
constexpr double CH_PI_2 = CH_PI / 2;

class ChVector3d {
  public:
    ChVector3d() : d_{0, 0, 0} {}
    ChVector3d(double x, double y, double z) : d_{x, y, z} {}
    double& x() { return d_[0]; }
    double& y() { return d_[1]; }
    double& z() { return d_[2]; }
    double x() const { return d_[0]; }
    double y() const { return d_[1]; }
    double z() const { return d_[2]; }
    double& operator[](int i) { return d_[i]; }
    double operator[](int i) const { return d_[i]; }
    double Length() const { return std::sqrt(d_[0] * d_[0] + d_[1] * d_[1] + d_[2] * d_[2]); }
    ChVector3d operator-() const { return ChVector3d(-d_[0], -d_[1], -d_[2]); }
    ChVector3d operator+(const ChVector3d& o) const { return ChVector3d(d_[0] + o.d_[0], d_[1] + o.d_[1], d_[2] + o.d_[2]); }
    ChVector3d operator-(const ChVector3d& o) const { return ChVector3d(d_[0] - o.d_[0], d_[1] - o.d_[1], d_[2] - o.d_[2]); }
    ChVector3d operator*(double s) const { return ChVector3d(d_[0] * s, d_[1] * s, d_[2] * s); }
    ChVector3d& operator+=(const ChVector3d& o) { d_[0] += o.d_[0]; d_[1] += o.d_[1]; d_[2] += o.d_[2]; return *this; }
    // cross product, Chrono's operator%
    ChVector3d operator%(const ChVector3d& o) const {
        return ChVector3d(d_[1] * o.d_[2] - d_[2] * o.d_[1], d_[2] * o.d_[0] - d_[0] * o.d_[2], d_[0] * o.d_[1] - d_[1] * o.d_[0]);
    }
    double Dot(const ChVector3d& o) const { return d_[0] * o.d_[0] + d_[1] * o.d_[1] + d_[2] * o.d_[2]; }

  private:
    double d_[3];
};
inline ChVector3d operator*(double s, const ChVector3d& v) { return v * s; }
inline ChVector3d Vcross(const ChVector3d& a, const ChVector3d& b) { return a % b; }

class ChQuaterniond {
  public:
    ChQuaterniond() : e0(1), e1(0), e2(0), e3(0) {}
    ChQuaterniond(double a, double b, double c, double d) : e0(a), e1(b), e2(c), e3(d) {}
    double e0, e1, e2, e3;
    void Normalize() {
        const double n = std::sqrt(e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3);
        if (n > 0) { e0 /= n; e1 /= n; e2 /= n; e3 /= n; }
    }
    ChQuaterniond operator*(const ChQuaterniond& q) const {
        return ChQuaterniond(e0 * q.e0 - e1 * q.e1 - e2 * q.e2 - e3 * q.e3, e0 * q.e1 + e1 * q.e0 + e2 * q.e3 - e3 * q.e2,
                             e0 * q.e2 - e1 * q.e3 + e2 * q.e0 + e3 * q.e1, e0 * q.e3 + e1 * q.e2 - e2 * q.e1 + e3 * q.e0);
    }
    // Cardan angles of the sequence R = Rx(a) Ry(b) Rz(c)
    ChVector3d GetCardanAnglesXYZ() const {
        const double r00 = 1 - 2 * (e2 * e2 + e3 * e3), r01 = 2 * (e1 * e2 - e0 * e3), r02 = 2 * (e1 * e3 + e0 * e2);
        const double r12 = 2 * (e2 * e3 - e0 * e1), r22 = 1 - 2 * (e1 * e1 + e2 * e2);
        double s = r02 > 1 ? 1 : (r02 < -1 ? -1 : r02);
        return ChVector3d(std::atan2(-r12, r22), std::asin(s), std::atan2(-r01, r00));
    }
    ChVector3d Rotate(const ChVector3d& v) const {
        const ChVector3d u(e1, e2, e3);
        const ChVector3d t = 2.0 * (u % v);
        return v + e0 * t + (u % t);
    }
    ChQuaterniond GetConjugate() const { return ChQuaterniond(e0, -e1, -e2, -e3); }
    ChVector3d RotateBack(const ChVector3d& v) const { return GetConjugate().Rotate(v); }
};
inline ChQuaterniond QuatFromAngleAxis(double angle, const ChVector3d& axis) {
    const double h = 0.5 * angle, s = std::sin(h);
    return ChQuaterniond(std::cos(h), axis.x() * s, axis.y() * s, axis.z() * s);
}
inline ChQuaterniond QuatFromAngleX(double a) { return QuatFromAngleAxis(a, ChVector3d(1, 0, 0)); }
inline ChQuaterniond QuatFromAngleY(double a) { return QuatFromAngleAxis(a, ChVector3d(0, 1, 0)); }
inline ChQuaterniond QuatFromAngleZ(double a) { return QuatFromAngleAxis(a, ChVector3d(0, 0, 1)); }

struct ChFramed {
    ChVector3d pos;
    ChQuaterniond rot;
    ChFramed() = default;
    explicit ChFramed(const ChVector3d& p, const ChQuaterniond& q = ChQuaterniond()) : pos(p), rot(q) {}
};

// Dense dynamic matrix / vector (row-major) standing in for ChMatrixDynamic<> / ChVectorDynamic<>.
template <class T = double>
class ChMatrixDynamic {
  public:
    ChMatrixDynamic() = default;
    ChMatrixDynamic(int r, int c) { setZero(r, c); }
    void setZero(int r, int c) { r_ = r; c_ = c; v_.assign(size_t(r) * c, T(0)); }
    void setZero() { v_.assign(v_.size(), T(0)); }
    void resize(int r, int c) { setZero(r, c); }
    int rows() const { return r_; }
    int cols() const { return c_; }
    T& operator()(int i, int j) { return v_[size_t(i) * c_ + j]; }
    const T& operator()(int i, int j) const { return v_[size_t(i) * c_ + j]; }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }

  private:
    int r_ = 0, c_ = 0;
    std::vector<T> v_;
};
template <class T = double>
class ChVectorDynamic {
  public:
    ChVectorDynamic() = default;
    explicit ChVectorDynamic(int n) : v_(size_t(n), T(0)) {}
    void setZero(int n) { v_.assign(size_t(n), T(0)); }
    void setZero() { v_.assign(v_.size(), T(0)); }
    int size() const { return int(v_.size()); }
    T& operator()(int i) { return v_[size_t(i)]; }
    const T& operator()(int i) const { return v_[size_t(i)]; }
    T& operator[](int i) { return v_[size_t(i)]; }
    const T& operator[](int i) const { return v_[size_t(i)]; }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }

  private:
    std::vector<T> v_;
};
using ChState = ChVectorDynamic<double>;
using ChStateDelta = ChVectorDynamic<double>;

class ChFunction {
  public:
    virtual ~ChFunction() = default;
    virtual ChFunction* Clone() const = 0;
    virtual double GetVal(double x) const = 0;
};

class ChSystem;
class ChBody;

class ChForce {
  public:
    enum class ForceType { FORCE, TORQUE };
    enum class AlignmentFrame { BODY_DIR, WORLD_DIR };
    void SetMode(ForceType m) { mode_ = m; }
    ForceType GetMode() const { return mode_; }
    void SetAlign(AlignmentFrame a) { align_ = a; }
    void SetName(const std::string& n) { name_ = n; }
    const std::string& GetName() const { return name_; }
    void SetF_x(std::shared_ptr<ChFunction> f) { f_[0] = std::move(f); }
    void SetF_y(std::shared_ptr<ChFunction> f) { f_[1] = std::move(f); }
    void SetF_z(std::shared_ptr<ChFunction> f) { f_[2] = std::move(f); }
    // ChForce::Update(time): evaluates the three component functions
    ChVector3d Evaluate(double t) const {
        return ChVector3d(f_[0] ? f_[0]->GetVal(t) : 0.0, f_[1] ? f_[1]->GetVal(t) : 0.0, f_[2] ? f_[2]->GetVal(t) : 0.0);
    }

  private:
    ForceType mode_ = ForceType::FORCE;
    AlignmentFrame align_ = AlignmentFrame::BODY_DIR;
    std::string name_;
    std::shared_ptr<ChFunction> f_[3];
};

class ChLoadable {
  public:
    virtual ~ChLoadable() = default;
};

class ChBody : public ChLoadable {
  public:
    void SetName(const std::string& n) { name_ = n; }
    const std::string& GetName() const { return name_; }
    void SetTag(int t) { tag_ = t; }
    void SetFixed(bool f) { fixed_ = f; }
    bool IsFixed() const { return fixed_; }
    void EnableCollision(bool) {}
    void SetMass(double m) { mass_ = m; }
    double GetMass() const { return mass_; }
    void SetInertiaXX(const ChVector3d& j) { inertia_ = j; }
    const ChVector3d& GetInertiaXX() const { return inertia_; }
    void SetPos(const ChVector3d& p) { pos_ = p; }
    const ChVector3d& GetPos() const { return pos_; }
    void SetRot(const ChQuaterniond& q) { rot_ = q; }
    const ChQuaterniond& GetRot() const { return rot_; }
    void SetPosDt(const ChVector3d& v) { vel_ = v; }
    const ChVector3d& GetPosDt() const { return vel_; }
    const ChVector3d& GetPosDt2() const { return acc_; }
    void SetAngVelParent(const ChVector3d& w) { wvel_ = w; }
    const ChVector3d& GetAngVelParent() const { return wvel_; }
    void AddForce(std::shared_ptr<ChForce> f) { forces_.push_back(std::move(f)); }
    const std::vector<std::shared_ptr<ChForce>>& GetForces() const { return forces_; }
    ChSystem* GetSystem() const { return system_; }
    double GetChTime() const;
    // compat-only: which of the 6 DoF are free (ChLinkLockPrismatic to a fixed body locks all but heave)
    bool free_dof[6] = {true, true, true, true, true, true};

  private:
    friend class ChSystem;
    std::string name_;
    int tag_ = 0;
    bool fixed_ = false;
    double mass_ = 1.0;
    ChVector3d inertia_{1, 1, 1};
    ChVector3d pos_, vel_, wvel_, acc_;
    ChQuaterniond rot_;
    std::vector<std::shared_ptr<ChForce>> forces_;
    ChSystem* system_ = nullptr;
};

// ChBodyEasyMesh(file, density, compute_mass, visualize, collide): geometry is irrelevant to the hydro path.
class ChBodyEasyMesh : public ChBody {
  public:
    ChBodyEasyMesh(const std::string& /*mesh*/, double /*density*/, bool = false, bool = false, bool = false) {}
};

struct ChLoadJacobians {
    ChMatrixDynamic<double> K, R, M;
};

class ChLoadBase {
  public:
    virtual ~ChLoadBase() = default;
    virtual void ComputeQ(ChState* state_x, ChStateDelta* state_w) = 0;
    virtual void ComputeJacobian(ChState* state_x, ChStateDelta* state_w) = 0;
    virtual void LoadIntLoadResidual_Mv(ChVectorDynamic<>& R, const ChVectorDynamic<>& w, const double c) = 0;
    virtual bool IsStiff() = 0;
    ChLoadJacobians* GetJacobians() { return m_jacobians.get(); }
    void CreateJacobianMatrices(int n) {
        m_jacobians = std::make_unique<ChLoadJacobians>();
        m_jacobians->K.setZero(n, n); m_jacobians->R.setZero(n, n); m_jacobians->M.setZero(n, n);
    }

  protected:
    std::unique_ptr<ChLoadJacobians> m_jacobians;
};

class ChLoadCustomMultiple : public ChLoadBase {
  public:
    explicit ChLoadCustomMultiple(std::vector<std::shared_ptr<ChLoadable>>& loadables) : loadables(loadables) {}
    virtual ChLoadCustomMultiple* Clone() const = 0;
    std::vector<std::shared_ptr<ChLoadable>> loadables;
};

class ChLoadContainer {
  public:
    void Add(std::shared_ptr<ChLoadBase> l) { loads_.push_back(std::move(l)); }
    const std::vector<std::shared_ptr<ChLoadBase>>& GetLoads() const { return loads_; }

  private:
    std::vector<std::shared_ptr<ChLoadBase>> loads_;
};

class ChLinkBase {
  public:
    virtual ~ChLinkBase() = default;
    void SetName(const std::string& n) { name_ = n; }
    const std::string& GetName() const { return name_; }
    virtual ChBody* GetBody1() const { return nullptr; }
    virtual ChBody* GetBody2() const { return nullptr; }

  private:
    std::string name_;
};
// Prismatic joint: the only relative motion left between the two bodies is a translation along the z axis of the
// joint frame (Chrono's convention).  One body fixed: the other keeps one DoF.  Both moving: body 1 becomes the
// 1-DoF child of body 2 (ChSystem handles it in reduced coordinates, chrono_compat.cpp).
class ChLinkLockPrismatic : public ChLinkBase {
  public:
    void Initialize(std::shared_ptr<ChBody> b1, std::shared_ptr<ChBody> b2, bool, const ChFramed& f1, const ChFramed&) { Set(b1, b2, f1); }
    void Initialize(std::shared_ptr<ChBody> b1, std::shared_ptr<ChBody> b2, const ChFramed& f) { Set(b1, b2, f); }
    ChBody* GetBody1() const override { return body1.get(); }
    ChBody* GetBody2() const override { return body2.get(); }
    // reaction of the joint on body 1 (world frame, at body 1's reference point) at the last force balance of a step;
    // body 2 receives the opposite force (and the torque transported to its own reference point)
    const ChVector3d& GetReactForce1() const { return react_force1; }
    const ChVector3d& GetReactTorque1() const { return react_torque1; }
    std::shared_ptr<ChBody> body1, body2;
    mutable ChVector3d react_force1, react_torque1;
    ChVector3d axis_in_parent{0, 0, 1};   // joint axis in body 2's frame (world frame when body 2 is fixed)
    ChVector3d offset_in_parent;          // body 1's position relative to body 2 at assembly, in body 2's frame
    ChQuaterniond rel_rot;                // body 1's orientation relative to body 2 at assembly

  private:
    void Set(const std::shared_ptr<ChBody>& b1, const std::shared_ptr<ChBody>& b2, const ChFramed& frame) {
        body1 = b1; body2 = b2;
        if (b1->IsFixed() && !b2->IsFixed()) std::swap(body1, body2);      // the moving body is the child
        const ChQuaterniond qp = body2->GetRot();
        axis_in_parent = qp.RotateBack(frame.rot.Rotate(ChVector3d(0, 0, 1)));
        offset_in_parent = qp.RotateBack(body1->GetPos() - body2->GetPos());
        rel_rot = qp.GetConjugate() * body1->GetRot();
    }
};
// Translational spring-damper between two points given in the absolute frame.
class ChLinkTSDA : public ChLinkBase {
  public:
    void Initialize(std::shared_ptr<ChBody> b1, std::shared_ptr<ChBody> b2, bool /*local*/, const ChVector3d& p1,
                    const ChVector3d& p2) {
        body1 = std::move(b1); body2 = std::move(b2);
        off1 = p1 - body1->GetPos(); off2 = p2 - body2->GetPos();
        loc1 = body1->GetRot().RotateBack(off1); loc2 = body2->GetRot().RotateBack(off2);
        rest = (p1 - p2).Length();
    }
    void SetSpringCoefficient(double k) { k_ = k; }
    void SetDampingCoefficient(double c) { c_ = c; }
    void SetRestLength(double r) { rest = r; }
    double GetSpringCoefficient() const { return k_; }
    double GetDampingCoefficient() const { return c_; }
    double GetRestLength() const { return rest; }
    ChBody* GetBody1() const override { return body1.get(); }
    ChBody* GetBody2() const override { return body2.get(); }
    ChVector3d GetPoint1Abs() const { return body1->GetPos() + body1->GetRot().Rotate(loc1); }
    ChVector3d GetPoint2Abs() const { return body2->GetPos() + body2->GetRot().Rotate(loc2); }
    double GetLength() const { return (GetPoint1Abs() - GetPoint2Abs()).Length(); }
    double GetVelocity() const {     // rate of change of the length
        const ChVector3d d = GetPoint1Abs() - GetPoint2Abs();
        const double len = d.Length();
        if (len == 0.0) return 0.0;
        const ChVector3d v1 = body1->GetPosDt() + (body1->GetAngVelParent() % body1->GetRot().Rotate(loc1));
        const ChVector3d v2 = body2->GetPosDt() + (body2->GetAngVelParent() % body2->GetRot().Rotate(loc2));
        return (v1 - v2).Dot(d * (1.0 / len));
    }
    double GetForce() const { return -(k_ * (GetLength() - rest) + c_ * GetVelocity()); }   // Chrono's sign: > 0 pushes apart
    std::shared_ptr<ChBody> body1, body2;
    ChVector3d off1, off2;      // attachment points relative to the bodies at assembly, world frame
    ChVector3d loc1, loc2;      // the same in the bodies' own frames
    double rest = 0, k_ = 0, c_ = 0;
};

struct ChSolver {
    enum class Type { GMRES, PSOR, BARZILAIBORWEIN, MINRES, SPARSE_LU, SPARSE_QR };
    struct Iterative { void SetMaxIterations(int) {} };
    Iterative* AsIterative() { return &it_; }
    Iterative it_;
};
struct ChTimestepper {
    enum class Type { EULER_IMPLICIT_LINEARIZED, HHT };
};

class ChSystem {
  public:
    virtual ~ChSystem() = default;
    void SetGravitationalAcceleration(const ChVector3d& g) { g_ = g; }
    const ChVector3d& GetGravitationalAcceleration() const { return g_; }
    void SetSolverType(ChSolver::Type) {}
    void SetTimestepperType(ChTimestepper::Type t) { stepper_ = t; }
    ChTimestepper::Type GetTimestepperType() const { return stepper_; }
    ChSolver* GetSolver() { return &solver_; }
    void Add(std::shared_ptr<ChBody> b) { AddBody(std::move(b)); }
    void AddBody(std::shared_ptr<ChBody> b) { b->system_ = this; bodies_.push_back(std::move(b)); }
    void Add(std::shared_ptr<ChLoadContainer> c) { load_containers_.push_back(std::move(c)); }
    void AddLink(std::shared_ptr<ChLinkBase> l) {
        if (auto t = std::dynamic_pointer_cast<ChLinkTSDA>(l)) tsdas_.push_back(t);
        if (auto p = std::dynamic_pointer_cast<ChLinkLockPrismatic>(l)) prismatics_.push_back(p);
        links_.push_back(std::move(l));
    }
    const std::vector<std::shared_ptr<ChLinkBase>>& GetLinks() const { return links_; }
    const std::vector<std::shared_ptr<ChLinkTSDA>>& GetTSDAs() const { return tsdas_; }
    const std::vector<std::shared_ptr<ChLinkLockPrismatic>>& GetPrismatics() const { return prismatics_; }
    double GetChTime() const { return time_; }
    void SetChTime(double t) { time_ = t; }
    double GetStep() const { return step_; }
    const std::vector<std::shared_ptr<ChBody>>& GetBodies() const { return bodies_; }
    // number of velocity-level coordinates = 6 per non-fixed body
    int GetNumCoordsVelLevel() const {
        int n = 0;
        for (auto& b : bodies_) if (!b->IsFixed()) n += 6;
        return n;
    }
    // Advances one step (see the header comment).  Returns 1 like ChSystem::DoStepDynamics.
    int DoStepDynamics(double dt);
    // The same step in two halves, for several systems advanced in lock-step around ONE batched force evaluation
    // (hydroc/hydro_ensemble.h): StepBegin brings the system to the state its forces are evaluated at (HHT: the
    // predictor at t + dt; linearised Euler: nothing, its forces belong to t), StepEnd evaluates the forces and
    // completes the step.  DoStepDynamics(dt) == StepBegin(dt); StepEnd().
    void StepBegin(double dt);
    void StepEnd();
    // compat-only: every system's StepBegin, then every system's StepEnd
    static void DoStepDynamicsLockstep(const std::vector<ChSystem*>& systems, double dt) {
        for (ChSystem* s : systems) s->StepBegin(dt);
        for (ChSystem* s : systems) s->StepEnd();
    }

  private:
    void Assemble(const std::vector<ChBody*>& act, std::vector<double>& F, std::vector<double>& M);
    std::vector<double> SolveAccelerations(const std::vector<ChBody*>& act, const std::vector<double>& rhs,
                                           const std::vector<double>& M);
    void ProjectOntoJoints(const std::vector<ChBody*>& act);
    void StepHHTBegin(const std::vector<ChBody*>& act, double h);
    void StepHHTEnd(const std::vector<ChBody*>& act, double h);
    std::vector<ChBody*> ActiveBodies() const;
    struct HHTSaved { ChVector3d x, v, w; ChQuaterniond q; };
    std::vector<HHTSaved> hht_s0_;
    double pending_dt_ = 0.0;
    ChTimestepper::Type stepper_ = ChTimestepper::Type::EULER_IMPLICIT_LINEARIZED;
    std::vector<double> hht_F_, hht_a_;      // generalised force and accelerations at t_n (HHT)
    ChVector3d g_{0, 0, -9.81};
    ChSolver solver_;
    double time_ = 0.0, step_ = 0.0;
    std::vector<std::shared_ptr<ChBody>> bodies_;
    std::vector<std::shared_ptr<ChLoadContainer>> load_containers_;
    std::vector<std::shared_ptr<ChLinkBase>> links_;
    std::vector<std::shared_ptr<ChLinkTSDA>> tsdas_;
    std::vector<std::shared_ptr<ChLinkLockPrismatic>> prismatics_;
};
class ChSystemNSC : public ChSystem {};
class ChSystemSMC : public ChSystem {};

inline double ChBody::GetChTime() const { return system_ ? system_->GetChTime() : 0.0; }

struct ChRealtimeStepTimer {};
inline void SetChronoDataPath(const std::string&) {}

}  // namespace chrono

#define CHRONO_VERSION "chrono_compat (stand-in)"
#define CHRONO_DATA_DIR ""
#endif
