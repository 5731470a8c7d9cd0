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

int ChSystem::DoStepDynamics(double dt) {
    step_ = dt;
    std::vector<ChBody*> act;
    for (auto& b : bodies_)
        if (!b->IsFixed()) act.push_back(b.get());
    const int n = 6 * int(act.size());
    if (n == 0) { time_ += dt; return 1; }

    // forces at (t_n, x_n, v_n): applied ChForces (world-aligned), gravity, spring-dampers
    std::vector<double> F(n, 0.0);
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

    // mass matrix: rigid-body inertia + the M Jacobians of stiff loads (ChLoadAddedMass)
    std::vector<double> M(size_t(n) * n, 0.0);
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

    // eliminate locked DoFs, solve for the accelerations of the free ones
    std::vector<int> idx;
    for (size_t i = 0; i < act.size(); ++i)
        for (int k = 0; k < 6; ++k)
            if (act[i]->free_dof[k]) idx.push_back(6 * int(i) + k);
    const int nf = int(idx.size());
    std::vector<double> A(size_t(nf) * nf), rhs(nf);
    for (int r = 0; r < nf; ++r) {
        rhs[r] = F[idx[r]];
        for (int c = 0; c < nf; ++c) A[size_t(r) * nf + c] = M[size_t(idx[r]) * n + idx[c]];
    }
    if (nf > 0) solve_dense(A, rhs, nf);
    std::vector<double> acc(n, 0.0);
    for (int r = 0; r < nf; ++r) acc[idx[r]] = rhs[r];

    // v_{n+1} = v_n + dt a ; x_{n+1} = x_n + dt v_{n+1}
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
        const double wl = w.Length();
        if (wl > 0.0) {
            ChQuaterniond q = QuatFromAngleAxis(wl * dt, w * (1.0 / wl)) * b->GetRot();
            q.Normalize();
            b->SetRot(q);
        }
    }
    time_ += dt;
    return 1;
}

}  // namespace chrono
#endif
