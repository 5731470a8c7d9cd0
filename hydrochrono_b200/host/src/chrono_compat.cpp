// Stepper of the chrono_compat stand-in (see chrono_compat.h).  Not used when building against real Chrono.
#ifndef HYDROC_HAVE_CHRONO
#include <chrono_compat/chrono_compat.h>

#include <algorithm>
#include <stdexcept>

namespace chrono {

static void solve_dense(std::vector<double>& A, std::vector<double>& b, int n) {
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int r = k + 1; r < n; ++r)
            if (std::fabs(A[size_t(r) * n + k]) > std::fabs(A[size_t(piv) * n + k])) piv = r;
        if (A[size_t(piv) * n + k] == 0.0) throw std::runtime_error("chrono_compat: singular mass matrix");
        if (piv != k) {
            for (int c = 0; c < n; ++c) std::swap(A[size_t(k) * n + c], A[size_t(piv) * n + c]);
            std::swap(b[k], b[piv]);
        }
        for (int r = k + 1; r < n; ++r) {
            const double m = A[size_t(r) * n + k] / A[size_t(k) * n + k];
            if (m == 0.0) continue;
            for (int c = k; c < n; ++c) A[size_t(r) * n + c] -= m * A[size_t(k) * n + c];
            b[r] -= m * b[k];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < n; ++c) s -= A[size_t(k) * n + c] * b[c];
        b[k] = s / A[size_t(k) * n + k];
    }
}

// Generalised force (applied ChForces evaluated at the current time/state, gravity, spring-dampers) and mass matrix
// (rigid-body inertia + the M Jacobians of stiff loads, i.e. ChLoadAddedMass) of the non-fixed bodies.
void ChSystem::Assemble(const std::vector<ChBody*>& act, std::vector<double>& F, std::vector<double>& M) {
    const int n = 6 * int(act.size());
    F.assign(n, 0.0);
    for (size_t i = 0; i < act.size(); ++i) {
        ChBody* b = act[i];
        for (const auto& f : b->GetForces()) {
            const ChVector3d v = f->Evaluate(time_);
            const int off = 6 * int(i) + (f->GetMode() == ChForce::ForceType::TORQUE ? 3 : 0);
            for (int k = 0; k < 3; ++k) F[off + k] += v[k];
        }
        for (int k = 0; k < 3; ++k) F[6 * i + k] += b->GetMass() * g_[k];
    }
    for (const auto& t : tsdas_) {
        // attachment points move with their bodies; the force acts along the line between them
        const ChVector3d s1 = t->body1->GetRot().Rotate(t->loc1), s2 = t->body2->GetRot().Rotate(t->loc2);
        const ChVector3d d = (t->body1->GetPos() + s1) - (t->body2->GetPos() + s2);
        const double len = d.Length();
        if (len == 0.0) continue;
        const ChVector3d u = d * (1.0 / len);
        const double f = t->GetForce();
        for (size_t i = 0; i < act.size(); ++i) {
            const bool one = act[i] == t->body1.get(), two = act[i] == t->body2.get();
            if (!one && !two) continue;
            const ChVector3d fv = u * (one ? f : -f), tq = (one ? s1 : s2) % fv;
            for (int k = 0; k < 3; ++k) { F[6 * i + k] += fv[k]; F[6 * i + 3 + k] += tq[k]; }
        }
    }
    M.assign(size_t(n) * n, 0.0);
    for (size_t i = 0; i < act.size(); ++i) {
        for (int k = 0; k < 3; ++k) {
            M[size_t(6 * i + k) * n + 6 * i + k] = act[i]->GetMass();
            M[size_t(6 * i + 3 + k) * n + 6 * i + 3 + k] = act[i]->GetInertiaXX()[k];
        }
    }
    for (auto& c : load_containers_)
        for (auto& l : c->GetLoads()) {
            if (!l->IsStiff()) continue;
            if (!l->GetJacobians() || l->GetJacobians()->M.rows() != n) l->CreateJacobianMatrices(n);
            l->ComputeJacobian(nullptr, nullptr);
            const auto& J = l->GetJacobians()->M;
            const int m = std::min(n, J.rows());
            for (int r = 0; r < m; ++r)
                for (int cc = 0; cc < m; ++cc) M[size_t(r) * n + cc] += J(r, cc);
        }
}

// Accelerations from M a = rhs in reduced coordinates.  Velocities of the active bodies are v = T qd with
//   root body:   its own 6 columns (those of locked DoFs dropped),
//   child body A of a prismatic joint to a moving parent B:  vA = vB + wB x rho + sd n,  wA = wB   (rho = xA - xB, n = joint axis)
//   body on a prismatic joint to a fixed body:  vA = sd n,  wA = 0,
// so  a = T qdd + bias  with  bias_A = wB x (wB x rho) + 2 sd wB x n  (Td qd), and
//   (T^T M T) qdd = T^T (rhs - M bias).
std::vector<double> ChSystem::SolveAccelerations(const std::vector<ChBody*>& act, const std::vector<double>& rhs_full,
                                                 const std::vector<double>& M) {
    const int n = 6 * int(act.size());
    auto index_of = [&](const ChBody* b) { for (size_t i = 0; i < act.size(); ++i) if (act[i] == b) return int(i); return -1; };
    std::vector<const ChLinkLockPrismatic*> joint(act.size(), nullptr);
    for (const auto& p : prismatics_) {
        const int c = index_of(p->body1.get());
        if (c < 0) continue;                                            // both ends fixed
        if (joint[c]) throw std::runtime_error("chrono_compat: a body may be the child of one prismatic joint only");
        joint[c] = p.get();
    }
    // columns: root bodies first (free DoFs), then one slide per jointed body
    std::vector<int> col0(act.size(), -1);
    std::vector<std::vector<int>> root_cols(act.size());
    int nq = 0;
    for (size_t i = 0; i < act.size(); ++i)
        if (!joint[i])
            for (int k = 0; k < 6; ++k)
                if (act[i]->free_dof[k]) { root_cols[i].push_back(k); if (col0[i] < 0) col0[i] = nq; ++nq; }
    std::vector<int> slide_col(act.size(), -1);
    for (size_t i = 0; i < act.size(); ++i) if (joint[i]) slide_col[i] = nq++;
    std::vector<double> T(size_t(n) * nq, 0.0), bias(n, 0.0);
    for (size_t i = 0; i < act.size(); ++i) {
        if (!joint[i]) {
            for (size_t c = 0; c < root_cols[i].size(); ++c) T[size_t(6 * i + root_cols[i][c]) * nq + col0[i] + int(c)] = 1.0;
            continue;
        }
        const ChLinkLockPrismatic& J = *joint[i];
        const int pi = index_of(J.body2.get());
        if (pi >= 0 && joint[pi]) throw std::runtime_error("chrono_compat: chains of prismatic joints are not supported");
        const ChVector3d nax = J.body2->GetRot().Rotate(J.axis_in_parent);
        for (int k = 0; k < 3; ++k) T[size_t(6 * i + k) * nq + slide_col[i]] = nax[k];
        if (pi < 0) continue;                                           // parent fixed: the slide is the only DoF
        const ChBody* A = act[i];
        const ChBody* B = act[pi];
        const ChVector3d rho = A->GetPos() - B->GetPos(), wB = B->GetAngVelParent();
        const double sd = nax.Dot(A->GetPosDt() - B->GetPosDt() - (wB % rho));
        // vA = vB - [rho]x wB ;  wA = wB   (only through the parent's free DoFs)
        const double skew[3][3] = {{0, rho.z(), -rho.y()}, {-rho.z(), 0, rho.x()}, {rho.y(), -rho.x(), 0}};   // -[rho]x
        for (size_t c = 0; c < root_cols[pi].size(); ++c) {
            const int k = root_cols[pi][c], col = col0[pi] + int(c);
            if (k < 3) T[size_t(6 * i + k) * nq + col] = 1.0;
            else {
                for (int r = 0; r < 3; ++r) T[size_t(6 * i + r) * nq + col] = skew[r][k - 3];
                T[size_t(6 * i + k) * nq + col] = 1.0;
            }
        }
        const ChVector3d bl = (wB % (wB % rho)) + 2.0 * sd * (wB % nax);
        for (int k = 0; k < 3; ++k) bias[6 * i + k] = bl[k];
    }
    // reduced system
    std::vector<double> r(n);
    for (int i = 0; i < n; ++i) {
        double s = rhs_full[i];
        for (int j = 0; j < n; ++j) s -= M[size_t(i) * n + j] * bias[j];
        r[i] = s;
    }
    std::vector<double> MT(size_t(n) * nq, 0.0), A(size_t(nq) * nq, 0.0), rq(nq, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double m = M[size_t(i) * n + j];
            if (m == 0.0) continue;
            for (int c = 0; c < nq; ++c) MT[size_t(i) * nq + c] += m * T[size_t(j) * nq + c];
        }
    for (int a = 0; a < nq; ++a) {
        for (int i = 0; i < n; ++i) {
            const double t = T[size_t(i) * nq + a];
            if (t == 0.0) continue;
            rq[a] += t * r[i];
            for (int c = 0; c < nq; ++c) A[size_t(a) * nq + c] += t * MT[size_t(i) * nq + c];
        }
    }
    if (nq > 0) solve_dense(A, rq, nq);
    std::vector<double> acc(bias);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < nq; ++c) acc[i] += T[size_t(i) * nq + c] * rq[c];
    // joint reaction on the child body: what the joint must supply for the child to follow the constrained motion,
    // R = (M a)_child - (applied force)_child
    for (size_t i = 0; i < act.size(); ++i) {
        if (!joint[i]) continue;
        double R[6];
        for (int k = 0; k < 6; ++k) {
            double sacc = -rhs_full[6 * i + k];
            for (int j = 0; j < n; ++j) sacc += M[size_t(6 * i + k) * n + j] * acc[j];
            R[k] = sacc;
        }
        joint[i]->react_force1 = ChVector3d(R[0], R[1], R[2]);
        joint[i]->react_torque1 = ChVector3d(R[3], R[4], R[5]);
    }
    return acc;
}

// Coordinate projection after a position / velocity update: every jointed body is put back on its joint (orientation
// and transverse offset taken from the parent, the slide kept), so the constraint never drifts.
void ChSystem::ProjectOntoJoints(const std::vector<ChBody*>& act) {
    (void)act;
    for (const auto& p : prismatics_) {
        ChBody* A = p->body1.get();
        ChBody* B = p->body2.get();
        if (A->IsFixed()) continue;
        const ChQuaterniond qB = B->GetRot();
        const ChVector3d nax = qB.Rotate(p->axis_in_parent);
        const ChVector3d rho_b = qB.RotateBack(A->GetPos() - B->GetPos());
        const double s = (rho_b - p->offset_in_parent).Dot(p->axis_in_parent);
        const ChVector3d rho = qB.Rotate(p->offset_in_parent + p->axis_in_parent * s);
        ChQuaterniond qA = qB * p->rel_rot;
        qA.Normalize();
        const ChVector3d wB = B->IsFixed() ? ChVector3d() : B->GetAngVelParent();
        const ChVector3d vB = B->IsFixed() ? ChVector3d() : B->GetPosDt();
        const double sd = nax.Dot(A->GetPosDt() - vB - (wB % rho));
        A->SetRot(qA);
        A->SetPos(B->GetPos() + rho);
        A->SetAngVelParent(wB);
        A->SetPosDt(vB + (wB % rho) + nax * sd);
    }
}

static void rotate_by(ChBody* b, const ChVector3d& dtheta) {
    const double a = dtheta.Length();
    if (a > 0.0) {
        ChQuaterniond q = QuatFromAngleAxis(a, dtheta * (1.0 / a)) * b->GetRot();
        q.Normalize();
        b->SetRot(q);
    }
}

std::vector<ChBody*> ChSystem::ActiveBodies() const {
    std::vector<ChBody*> act;
    for (auto& b : bodies_)
        if (!b->IsFixed()) act.push_back(b.get());
    return act;
}

int ChSystem::DoStepDynamics(double dt) {
    StepBegin(dt);
    StepEnd();
    return 1;
}

void ChSystem::StepBegin(double dt) {
    step_ = dt;
    pending_dt_ = dt;
    const std::vector<ChBody*> act = ActiveBodies();
    if (!act.empty() && stepper_ == ChTimestepper::Type::HHT) StepHHTBegin(act, dt);
}

void ChSystem::StepEnd() {
    const double dt = pending_dt_;
    const std::vector<ChBody*> act = ActiveBodies();
    const int n = 6 * int(act.size());
    if (n == 0) { time_ += dt; return; }
    if (stepper_ == ChTimestepper::Type::HHT) { StepHHTEnd(act, dt); return; }

    // linearised Euler: forces at (t_n, x_n, v_n);  v_{n+1} = v_n + dt a ; x_{n+1} = x_n + dt v_{n+1}
    std::vector<double> F, M;
    Assemble(act, F, M);
    const std::vector<double> acc = SolveAccelerations(act, F, M);
    for (size_t i = 0; i < act.size(); ++i) {
        ChBody* b = act[i];
        ChVector3d v = b->GetPosDt(), w = b->GetAngVelParent();
        for (int k = 0; k < 3; ++k) {
            v[k] = b->free_dof[k] ? v[k] + dt * acc[6 * i + k] : 0.0;
            w[k] = b->free_dof[3 + k] ? w[k] + dt * acc[6 * i + 3 + k] : 0.0;
        }
        b->acc_ = ChVector3d(acc[6 * i], acc[6 * i + 1], acc[6 * i + 2]);
        b->SetPosDt(v);
        b->SetAngVelParent(w);
        b->SetPos(b->GetPos() + v * dt);
        rotate_by(b, w * dt);
    }
    ProjectOntoJoints(act);
    time_ += dt;
}

// HHT-alpha (alpha = -0.2, gamma = 1/2 - alpha, beta = (1 - alpha)^2 / 4), the stepper of the reference's YAML runs.
// The applied forces are time-keyed callbacks (ChFunction::GetVal(t); TestHydro caches per time value), so within a
// step they are evaluated exactly once, at t_{n+1} with the predictor state -- later iterations of Chrono's Newton
// loop would read the cache.  The step is therefore:
//   predictor   x* = x_n + h v_n + h^2/2 a_n,  v* = v_n + h a_n                       (a* = a_n)
//   forces      F_{n+1} = F(t_{n+1}, x*, v*)
//   balance     M a_{n+1} = (1 + alpha) F_{n+1} - alpha F_n
//   corrector   x_{n+1} = x_n + h v_n + h^2 ((1/2 - beta) a_n + beta a_{n+1}),  v_{n+1} = v_n + h ((1 - gamma) a_n + gamma a_{n+1})
void ChSystem::StepHHTBegin(const std::vector<ChBody*>& act, double h) {
    const int n = 6 * int(act.size());
    std::vector<double> F, M;
    if (int(hht_F_.size()) != n) {                       // first step: consistent initial accelerations M a_0 = F_0
        Assemble(act, F, M);
        hht_F_ = F;
        hht_a_ = SolveAccelerations(act, F, M);
    }
    std::vector<HHTSaved>& s0 = hht_s0_;
    s0.resize(act.size());
    for (size_t i = 0; i < act.size(); ++i) {
        ChBody* b = act[i];
        s0[i] = HHTSaved{b->GetPos(), b->GetPosDt(), b->GetAngVelParent(), b->GetRot()};
        const ChVector3d a(hht_a_[6 * i], hht_a_[6 * i + 1], hht_a_[6 * i + 2]);
        const ChVector3d al(hht_a_[6 * i + 3], hht_a_[6 * i + 4], hht_a_[6 * i + 5]);
        b->SetPos(s0[i].x + s0[i].v * h + a * (0.5 * h * h));
        b->SetPosDt(s0[i].v + a * h);
        rotate_by(b, s0[i].w * h + al * (0.5 * h * h));
        b->SetAngVelParent(s0[i].w + al * h);
    }
    ProjectOntoJoints(act);
    time_ += h;
}

void ChSystem::StepHHTEnd(const std::vector<ChBody*>& act, double h) {
    const double alpha = -0.2, gamma = 0.5 - alpha, beta = 0.25 * (1.0 - alpha) * (1.0 - alpha);
    const int n = 6 * int(act.size());
    const std::vector<HHTSaved>& s0 = hht_s0_;
    std::vector<double> F, M;
    Assemble(act, F, M);
    std::vector<double> rhs(n);
    for (int k = 0; k < n; ++k) rhs[k] = (1.0 + alpha) * F[k] - alpha * hht_F_[k];
    const std::vector<double> an = SolveAccelerations(act, rhs, M);
    for (size_t i = 0; i < act.size(); ++i) {
        ChBody* b = act[i];
        ChVector3d x = s0[i].x, v = s0[i].v, w = s0[i].w, dth;
        for (int k = 0; k < 3; ++k) {
            const double a0 = hht_a_[6 * i + k], a1 = an[6 * i + k];
            const double l0 = hht_a_[6 * i + 3 + k], l1 = an[6 * i + 3 + k];
            if (b->free_dof[k]) {
                x[k] += h * s0[i].v[k] + h * h * ((0.5 - beta) * a0 + beta * a1);
                v[k] += h * ((1.0 - gamma) * a0 + gamma * a1);
            } else v[k] = 0.0;
            if (b->free_dof[3 + k]) {
                dth[k] = h * s0[i].w[k] + h * h * ((0.5 - beta) * l0 + beta * l1);
                w[k] += h * ((1.0 - gamma) * l0 + gamma * l1);
            } else { dth[k] = 0.0; w[k] = 0.0; }
        }
        b->SetPos(x);
        b->SetPosDt(v);
        b->SetRot(s0[i].q);
        rotate_by(b, dth);
        b->SetAngVelParent(w);
        b->acc_ = ChVector3d(an[6 * i], an[6 * i + 1], an[6 * i + 2]);
    }
    ProjectOntoJoints(act);
    hht_F_ = F;
    hht_a_ = an;
}

}  // namespace chrono
#endif
