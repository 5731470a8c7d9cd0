// ChLoadAddedMass (reference src/chloadaddedmass.cpp:12-71).
#include <hydroc/chloadaddedmass.h>

ChLoadAddedMass::ChLoadAddedMass(const std::vector<HydroData::BodyInfo>& body_info,
                                 std::vector<std::shared_ptr<ChLoadable>>& bodies, ChSystem* system)
    : ChLoadCustomMultiple(bodies), system(system) {
    const int nBodies = int(bodies.size());
    infinite_added_mass.setZero(6 * nBodies, 6 * nBodies);
    for (int i = 0; i < nBodies; i++)
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6 * nBodies; ++c) infinite_added_mass(6 * i + r, c) = body_info[i].inf_added_mass(r, c);
    infinite_added_mass_system = infinite_added_mass;
}

void ChLoadAddedMass::ComputeJacobian(ChState*, ChStateDelta*) {
    // pad to the size of the system mass matrix; the hydro block sits at (0,0), so hydro bodies must be the
    // first bodies added to the system (reference :35-44)
    const int mmrows = system->GetNumCoordsVelLevel();
    if (mmrows != infinite_added_mass_system.rows() && mmrows > 0) {
        infinite_added_mass_system.setZero(mmrows, mmrows);
        const int am = infinite_added_mass.rows();
        for (int r = 0; r < am && r < mmrows; ++r)
            for (int c = 0; c < am && c < mmrows; ++c) infinite_added_mass_system(r, c) = infinite_added_mass(r, c);
    }
    if (!m_jacobians) CreateJacobianMatrices(infinite_added_mass_system.rows());
    m_jacobians->M = infinite_added_mass_system;
    m_jacobians->R.setZero();
    m_jacobians->K.setZero();
}

void ChLoadAddedMass::LoadIntLoadResidual_Mv(ChVectorDynamic<>& R, const ChVectorDynamic<>& w, const double c) {
    if (!this->m_jacobians) return;
    const auto& M = m_jacobians->M;      // R += c * M * w
    for (int i = 0; i < M.rows(); ++i) {
        double s = 0.0;
        for (int j = 0; j < M.cols(); ++j) s += (c * M(i, j)) * w(j);
        R(i) += s;
    }
}
