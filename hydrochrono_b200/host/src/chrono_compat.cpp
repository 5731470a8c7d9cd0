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
        const ChVector3d p1 = t->body1->GetPos() + t->off1, p2 = t->body2->GetPos() + t->off2;
        const ChVector3d d = p1 - p2;
        const double len = d.Length();
        if (len == 0.0) continue;
        const ChVector3d u = d * (1.0 / len);
        const double vrel = (t->body1->GetPosDt() - t->body2->GetPosDt()).Dot(u);
        const double f = -(t->k_ * (len - t->rest) + t->c_ * vrel);
        for (size_t i = 0; i < act.size(); ++i) {
            const double sgn = act[i] == t->body1.get() ? 1.0 : (act[i] == t->body2.get() ? -1.0 : 0.0);
            for (int k = 0; k < 3; ++k) F[6 * i + k] += sgn * f * u[k];
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

// Accelerations of the free DoFs from M a = rhs (locked DoFs eliminated, their acceleration is zero).
std::vector<double> ChSystem::SolveAccelerations(const std::vector<ChBody*>& act, const std::vector<double>& rhs_full,
                                                 const std::vector<double>& M) {
    const int n = 6 * int(act.size());
    std::vector<int> idx;
    for (size_t i = 0; i < act.size(); ++i)
        for (int k = 0; k < 6; ++k)
            if (act[i]->free_dof[k]) idx.push_back(6 * int(i) + k);
    const int nf = int(idx.size());
    std::vector<double> A(size_t(nf) * nf), rhs(nf);
    for (int r = 0; r < nf; ++r) {
        rhs[r] = rhs_full[idx[r]];
        for (int c = 0; c < nf; ++c) A[size_t(r) * nf + c] = M[size_t(idx[r]) * n + idx[c]];
    }
    if (nf > 0) solve_dense(A, rhs, nf);
    std::vector<double> acc(n, 0.0);
    for (int r = 0; r < nf; ++r) acc[idx[r]] = rhs[r];
    return acc;
}

static void rotate_by(ChBody* b, const ChVector3d& dtheta) {
    const double a = dtheta.Length();
    if (a > 0.0) {
        ChQuaterniond q = QuatFromAngleAxis(a, dtheta * (1.0 / a)) * b->GetRot();
        q.Normalize();
        b->SetRot(q);
    }
}

int ChSystem::DoStepDynamics(double dt) {
    step_ = dt;
    std::vector<ChBody*> act;
    for (auto& b : bodies_)
        if (!b->IsFixed()) act.push_back(b.get());
    const int n = 6 * int(act.size());
    if (n == 0) { time_ += dt; return 1; }
    if (stepper_ == ChTimestepper::Type::HHT) return StepHHT(act, dt);

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
    time_ += dt;
    return 1;
}

// HHT-alpha (alpha = -0.2, gamma = 1/2 - alpha, beta = (1 - alpha)^2 / 4), the stepper of the reference's YAML runs.
// The applied forces are time-keyed callbacks (ChFunction::GetVal(t); TestHydro caches per time value), so within a
// step they are evaluated exactly once, at t_{n+1} with the predictor state -- later iterations of Chrono's Newton
// loop would read the cache.  The step is therefore:
//   predictor   x* = x_n + h v_n + h^2/2 a_n,  v* = v_n + h a_n                       (a* = a_n)
//   forces      F_{n+1} = F(t_{n+1}, x*, v*)
//   balance     M a_{n+1} = (1 + alpha) F_{n+1} - alpha F_n
//   corrector   x_{n+1} = x_n + h v_n + h^2 ((1/2 - beta) a_n + beta a_{n+1}),  v_{n+1} = v_n + h ((1 - gamma) a_n + gamma a_{n+1})
int ChSystem::StepHHT(const std::vector<ChBody*>& act, double h) {
    const double alpha = -0.2, gamma = 0.5 - alpha, beta = 0.25 * (1.0 - alpha) * (1.0 - alpha);
    const int n = 6 * int(act.size());
    std::vector<double> F, M;
    if (int(hht_F_.size()) != n) {                       // first step: consistent initial accelerations M a_0 = F_0
        Assemble(act, F, M);
        hht_F_ = F;
        hht_a_ = SolveAccelerations(act, F, M);
    }
    struct Saved { ChVector3d x, v, w; ChQuaterniond q; };
    std::vector<Saved> s0(act.size());
    for (size_t i = 0; i < act.size(); ++i) {
        ChBody* b = act[i];
        s0[i] = Saved{b->GetPos(), b->GetPosDt(), b->GetAngVelParent(), b->GetRot()};
        const ChVector3d a(hht_a_[6 * i], hht_a_[6 * i + 1], hht_a_[6 * i + 2]);
        const ChVector3d al(hht_a_[6 * i + 3], hht_a_[6 * i + 4], hht_a_[6 * i + 5]);
        b->SetPos(s0[i].x + s0[i].v * h + a * (0.5 * h * h));
        b->SetPosDt(s0[i].v + a * h);
        rotate_by(b, s0[i].w * h + al * (0.5 * h * h));
        b->SetAngVelParent(s0[i].w + al * h);
    }
    time_ += h;
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
    hht_F_ = F;
    hht_a_ = an;
    return 1;
}

}  // namespace chrono
#endif
