// TestHydro, ForceFunc6d, ComponentFunc (reference src/hydro_forces.cpp) over the C ABI of libhydrochrono_b200.
//
// Chrono evaluates six ChFunctions per body (ComponentFunc::GetVal -> ForceFunc6d::CoordinateFunc ->
// TestHydro::CoordinateFuncForBody).  The first evaluation at a new ChTime gathers the state of every hydro
// body, runs ONE device step (hc_step: H2D, kernels, D2H) and caches the 6N totals; later evaluations at the
// same time return the cache -- the reference's own once-per-time-value contract (:742-755).
#include <hydroc/hydro_forces.h>

#include <cstdlib>
#include <iomanip>

#include <hydroc/chloadaddedmass.h>

#include <cstring>
#include <iostream>
#include <stdexcept>

#include "hc_check.h"

static const int kDofPerBody = 6;

// ---------------------------------------------------------------------------------------------
// ComponentFunc (:63-85)
// ---------------------------------------------------------------------------------------------
ComponentFunc::ComponentFunc() : base_(nullptr), index_(kDofPerBody) {}
ComponentFunc::ComponentFunc(ForceFunc6d* b, int i) : base_(b), index_(i) {}
ComponentFunc::ComponentFunc(const ComponentFunc& old) : ChFunction(old), base_(old.base_), index_(old.index_) {}
ComponentFunc* ComponentFunc::Clone() const { return new ComponentFunc(*this); }

double ComponentFunc::GetVal(double) const {
    if (base_ == nullptr) {
        std::cout << "base == Null!" << std::endl;
        return 0;
    }
    return base_->CoordinateFunc(index_);
}

// ---------------------------------------------------------------------------------------------
// ForceFunc6d (:87-168): two world-aligned ChForces (force, torque) whose components call back into TestHydro
// ---------------------------------------------------------------------------------------------
ForceFunc6d::ForceFunc6d()
    : b_num_(0), forces_{{this, 0}, {this, 1}, {this, 2}, {this, 3}, {this, 4}, {this, 5}}, all_hydro_forces_(nullptr) {
    for (unsigned i = 0; i < 6; i++)   // non-owning shared_ptrs to the member functors
        force_ptrs_[i] = std::shared_ptr<ComponentFunc>(forces_ + i, [](ComponentFunc*) {});
    chrono_force_ = chrono_types::make_shared<ChForce>();
    chrono_torque_ = chrono_types::make_shared<ChForce>();
    chrono_force_->SetAlign(ChForce::AlignmentFrame::WORLD_DIR);
    chrono_torque_->SetAlign(ChForce::AlignmentFrame::WORLD_DIR);
    chrono_force_->SetName("hydroforce");
    chrono_torque_->SetName("hydrotorque");
}

ForceFunc6d::ForceFunc6d(std::shared_ptr<ChBody> object, TestHydro* user_all_forces) : ForceFunc6d() {
    body_ = object;
    std::string temp = body_->GetName();   // "bodyN" -> N (1-indexed), reference :106-107
    b_num_ = stoi(temp.erase(0, 4));
    all_hydro_forces_ = user_all_forces;
    if (all_hydro_forces_ == nullptr) std::cout << "all hydro forces null " << std::endl;
    SetForce();
    SetTorque();
    ApplyForceAndTorqueToBody();
}

ForceFunc6d::ForceFunc6d(const ForceFunc6d& old)
    : forces_{{this, 0}, {this, 1}, {this, 2}, {this, 3}, {this, 4}, {this, 5}} {
    for (unsigned i = 0; i < 6; i++)
        force_ptrs_[i] = std::shared_ptr<ComponentFunc>(forces_ + i, [](ComponentFunc*) {});
    chrono_force_ = old.chrono_force_;
    chrono_torque_ = old.chrono_torque_;
    body_ = old.body_;
    b_num_ = old.b_num_;
    all_hydro_forces_ = old.all_hydro_forces_;
    SetForce();      // re-point the ChForces at this copy's functors; the forces are NOT added to the body again
    SetTorque();
}

double ForceFunc6d::CoordinateFunc(int i) {
    if (i >= kDofPerBody || i < 0) {
        std::cout << "wrong index force func 6d" << std::endl;
        return 0;
    }
    return all_hydro_forces_->CoordinateFuncForBody(b_num_, i);
}

void ForceFunc6d::SetForce() {
    if (chrono_force_ == nullptr || body_ == nullptr) std::cout << "set force null issue" << std::endl;
    chrono_force_->SetF_x(force_ptrs_[0]);
    chrono_force_->SetF_y(force_ptrs_[1]);
    chrono_force_->SetF_z(force_ptrs_[2]);
}
void ForceFunc6d::SetTorque() {
    if (chrono_torque_ == nullptr || body_ == nullptr) std::cout << "set torque null issue" << std::endl;
    chrono_torque_->SetF_x(force_ptrs_[3]);
    chrono_torque_->SetF_y(force_ptrs_[4]);
    chrono_torque_->SetF_z(force_ptrs_[5]);
    chrono_torque_->SetMode(ChForce::ForceType::TORQUE);
}
void ForceFunc6d::ApplyForceAndTorqueToBody() {
    body_->AddForce(chrono_force_);
    body_->AddForce(chrono_torque_);
}

// ---------------------------------------------------------------------------------------------
// TestHydro
// ---------------------------------------------------------------------------------------------
TestHydro::TestHydro(std::vector<std::shared_ptr<ChBody>> user_bodies, std::string h5_file_name,
                     std::shared_ptr<WaveBase> waves)
    : bodies_(user_bodies), num_bodies_(int(bodies_.size())),
      file_info_(H5FileInfo(h5_file_name, int(bodies_.size())).ReadH5Data()) {
    Construct(std::move(waves));
}

TestHydro::TestHydro(std::vector<std::shared_ptr<ChBody>> user_bodies, hc_tables* tables, std::shared_ptr<WaveBase> waves)
    : bodies_(user_bodies), num_bodies_(int(bodies_.size())), file_info_(H5FileInfo::FromTables(tables, "<memory>")) {
    if (hc_tables_num_bodies(tables) != num_bodies_)
        throw std::runtime_error("TestHydro: number of bodies does not match the hydro tables");
    Construct(std::move(waves));
}

void TestHydro::Construct(std::shared_ptr<WaveBase> waves) {
    prev_time = -1;   // reference :176
    if (bodies_.empty() || !bodies_[0]) throw std::runtime_error("bodies_ array is empty or invalid in TestHydro");
    const int total_dofs = kDofPerBody * num_bodies_;
    force_hydrostatic_.assign(total_dofs, 0.0);
    force_radiation_damping_.assign(total_dofs, 0.0);
    force_waves_.assign(total_dofs, 0.0);
    total_force_.assign(total_dofs, 0.0);

    // device ensemble for this one system (B = 1)
    hc_ensemble_opts o;
    hc_ensemble_default_opts(&o);
    o.batch = 1;
    ChSystem* sys = bodies_[0]->GetSystem();
    o.dt_hint = sys ? sys->GetStep() : 0.0;
    hc_throw_on_error(hc_ensemble_create(file_info_.handle(), &o, &ens_));

    for (int b = 0; b < num_bodies_; ++b) force_per_body_.emplace_back(bodies_[b], this);
    if (const char* tr = std::getenv("HYDROC_STATE_TRACE")) trace_ = std::make_unique<std::ofstream>(tr);

    // added mass (reference :223-234)
    my_loadcontainer = chrono_types::make_shared<ChLoadContainer>();
    std::vector<std::shared_ptr<ChLoadable>> loadables(bodies_.size());
    for (size_t i = 0; i < bodies_.size(); ++i) loadables[i] = bodies_[i];
    my_loadbodyinertia = chrono_types::make_shared<ChLoadAddedMass>(file_info_.GetBodyInfos(), loadables, sys);
    if (sys) sys->Add(my_loadcontainer);
    my_loadcontainer->Add(my_loadbodyinertia);

    AddWaves(std::move(waves));
}

TestHydro::~TestHydro() {
    if (ens_) hc_ensemble_destroy(ens_);
}

void TestHydro::AddWaves(std::shared_ptr<WaveBase> waves) {
    user_waves_ = std::move(waves);
    switch (user_waves_->GetWaveMode()) {
        case WaveMode::regular: {
            auto reg = std::static_pointer_cast<RegularWave>(user_waves_);
            reg->AddH5Data(file_info_.GetRegularWaveInfos(), file_info_.GetSimulationInfo());
            break;
        }
        case WaveMode::irregular: {
            auto irreg = std::static_pointer_cast<IrregularWaves>(user_waves_);
            irreg->AddH5Data(file_info_.GetIrregularWaveInfos(), file_info_.GetSimulationInfo());
            break;
        }
        default: break;
    }
    user_waves_->Bind(ens_, file_info_.GetSimulationInfo(), static_cast<unsigned int>(num_bodies_));
    user_waves_->Initialize();
}

void TestHydro::SetRadiationConvolutionMode(RadiationConvolutionMode mode) {
    convolution_mode_ = mode;
    convolution_dirty_ = true;
}
void TestHydro::SetTaperedDirectOptions(const TaperedDirectOptions& opts) {
    tapered_opts_ = opts;
    convolution_dirty_ = true;
}

// EnsureProcessedRIRF (:385-535): preprocessing happens lazily before the first use of the kernel
void TestHydro::ApplyConvolutionMode() {
    if (!convolution_dirty_) return;
    hc_tapered_opts o;
    o.smoothing = tapered_opts_.smoothing.c_str();
    o.window_length = tapered_opts_.window_length;
    o.rirf_end_time = tapered_opts_.rirf_end_time;
    o.taper_start_percent = tapered_opts_.taper_start_percent;
    o.taper_end_percent = tapered_opts_.taper_end_percent;
    o.taper_final_amplitude = tapered_opts_.taper_final_amplitude;
    const int mode = convolution_mode_ == RadiationConvolutionMode::TaperedDirect ? 1 : 0;
    hc_throw_on_error(hc_tables_set_convolution_mode(file_info_.handle(), mode, &o));
    hc_throw_on_error(hc_ensemble_refresh_rirf(ens_));
    convolution_dirty_ = false;
}

double TestHydro::GetRIRFval(int row, int col, int st) {
    ApplyConvolutionMode();
    double v = 0;
    hc_throw_on_error(hc_tables_rirf_val(file_info_.handle(), row, col, st, &v));
    return v;
}

// One device step at the current Chrono time with the state Chrono has scattered at this moment.
void TestHydro::EvaluateAtCurrentTime() {
    ApplyConvolutionMode();
    const int D = kDofPerBody * num_bodies_;
    std::vector<double> pose(D), vel(D);
    for (int b = 0; b < num_bodies_; ++b) {
        const auto& body = bodies_[b];
        const ChVector3d p = body->GetPos();
        const ChVector3d a = body->GetRot().GetCardanAnglesXYZ();   // reference :279-280
        const ChVector3d v = body->GetPosDt();
        const ChVector3d w = body->GetAngVelParent();               // reference :567-568
        double* x = &pose[kDofPerBody * b];
        double* u = &vel[kDofPerBody * b];
        x[0] = p.x(); x[1] = p.y(); x[2] = p.z(); x[3] = a.x(); x[4] = a.y(); x[5] = a.z();
        u[0] = v.x(); u[1] = v.y(); u[2] = v.z(); u[3] = w.x(); u[4] = w.y(); u[5] = w.z();
    }
    const ChVector3d g = bodies_[0]->GetSystem()->GetGravitationalAcceleration();
    const double gv[3] = {g.x(), g.y(), g.z()};
    const double t = bodies_[0]->GetChTime();
    hc_throw_on_error(hc_step(ens_, t, pose.data(), vel.data(), gv, total_force_.data(), nullptr));
    prev_time = t;
    components_fetched_ = false;
    if (trace_) {   // HYDROC_STATE_TRACE: one line per evaluation -- t, pose[6N], vel[6N], g[3], total force[6N]
        *trace_ << std::setprecision(17) << t;
        for (double x : pose) *trace_ << ' ' << x;
        for (double x : vel) *trace_ << ' ' << x;
        *trace_ << ' ' << gv[0] << ' ' << gv[1] << ' ' << gv[2];
        for (double x : total_force_) *trace_ << ' ' << x;
        *trace_ << '\n';
    }
}

static void fetch_components(hc_ensemble* ens, std::vector<double>& hs, std::vector<double>& rad, std::vector<double>& wv) {
    hc_throw_on_error(hc_get_components(ens, hs.data(), rad.data(), wv.data()));
}

std::vector<double> TestHydro::ComputeForceHydrostatics() {
    if (bodies_[0]->GetChTime() != prev_time) EvaluateAtCurrentTime();
    if (!components_fetched_) { fetch_components(ens_, force_hydrostatic_, force_radiation_damping_, force_waves_); components_fetched_ = true; }
    return force_hydrostatic_;
}

std::vector<double> TestHydro::ComputeForceRadiationDampingConv() {
    if (bodies_[0]->GetChTime() != prev_time) EvaluateAtCurrentTime();
    if (!components_fetched_) { fetch_components(ens_, force_hydrostatic_, force_radiation_damping_, force_waves_); components_fetched_ = true; }
    return force_radiation_damping_;
}

Eigen::VectorXd TestHydro::ComputeForceWaves() {
    if (bodies_.empty()) throw std::runtime_error("bodies_ array is empty in ComputeForceWaves");
    if (bodies_[0]->GetChTime() != prev_time) EvaluateAtCurrentTime();
    if (!components_fetched_) { fetch_components(ens_, force_hydrostatic_, force_radiation_damping_, force_waves_); components_fetched_ = true; }
    return Eigen::VectorXd(force_waves_);
}

double TestHydro::CoordinateFuncForBody(int b, int dof_index) {
    if (dof_index < 0 || dof_index >= kDofPerBody || b < 1 || b > num_bodies_)
        throw std::out_of_range("Invalid index in CoordinateFuncForBody");
    if (bodies_.empty() || !bodies_[0]) throw std::runtime_error("bodies_ array is empty or invalid in CoordinateFuncForBody");
    const int body_num_offset = kDofPerBody * (b - 1);
    if (bodies_[0]->GetChTime() != prev_time) EvaluateAtCurrentTime();   // else: cached totals (:742-744)
    return total_force_[body_num_offset + dof_index];
}

HydroProfileStats TestHydro::GetProfileStats() const {
    hc_profile_stats s;
    HydroProfileStats out;
    if (hc_get_profile(ens_, &s) == HC_OK) {
        out.hydrostatics_seconds = s.hydrostatics_seconds; out.radiation_seconds = s.radiation_seconds;
        out.waves_seconds = s.waves_seconds; out.hydrostatics_calls = s.hydrostatics_calls;
        out.radiation_calls = s.radiation_calls; out.waves_calls = s.waves_calls;
    }
    return out;
}
