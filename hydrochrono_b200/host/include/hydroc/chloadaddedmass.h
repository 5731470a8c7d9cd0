// ChLoadAddedMass with the reference's public surface (include/hydroc/chloadaddedmass.h:22-90): the
// infinite-frequency added mass as a stiff Chrono load.  The matrix comes from hc_added_mass().
#ifndef HYDROC_B200_CHLOADADDEDMASS_H
#define HYDROC_B200_CHLOADADDEDMASS_H
#pragma once

#include <vector>

#include <chrono_compat/chrono_compat.h>
#include <hydroc/h5fileinfo.h>

using namespace chrono;

class ChLoadAddedMass : public chrono::ChLoadCustomMultiple {
  public:
    // body_info_struct: HydroData::GetBodyInfos() (per-body 6 x 6N blocks, already x rho)
    ChLoadAddedMass(const std::vector<HydroData::BodyInfo>& body_info_struct,
                    std::vector<std::shared_ptr<ChLoadable>>& bodies, ChSystem* system);

    virtual ChLoadAddedMass* Clone() const override { return new ChLoadAddedMass(*this); }
    virtual void ComputeQ(ChState*, ChStateDelta*) override {}
    virtual void ComputeJacobian(ChState* state_x, ChStateDelta* state_w) override;
    virtual void LoadIntLoadResidual_Mv(ChVectorDynamic<>& R, const ChVectorDynamic<>& w, const double c) override;
    virtual bool IsStiff() override { return true; }   // forces the use of the M, R, K matrices

    ChLoadAddedMass(const ChLoadAddedMass& o)
        : ChLoadCustomMultiple(const_cast<std::vector<std::shared_ptr<ChLoadable>>&>(o.loadables)), system(o.system),
          infinite_added_mass(o.infinite_added_mass), infinite_added_mass_system(o.infinite_added_mass_system) {}

  private:
    ChSystem* system;
    ChMatrixDynamic<double> infinite_added_mass;         // 6N x 6N
    ChMatrixDynamic<double> infinite_added_mass_system;  // padded to the system's velocity-level size
};

#endif
