// Results export in HydroChrono's HDF5 schema v0.3 (reference src/simulation_exporter.cpp:181-199,303-391,723-748).
#include <hydroc/simulation_exporter.h>

#include <stdexcept>

#include "hc_check.h"

namespace hydroc {

struct SimulationExporter::Impl {
    Options options;
    hc_h5_writer* w = nullptr;
    bool finalized = false;
    struct BodyBuf {
        std::string name;
        std::vector<double> pos, vel, acc, quat, euler, wvel;
    };
    std::vector<BodyBuf> bodies;
    std::vector<double> time;
    // links (reference src/simulation_exporter.cpp:103-152): translational spring-dampers and joints
    struct TsdaBuf {
        std::string name;
        chrono::ChLinkTSDA* link = nullptr;
        double rest_length = 0, k = 0, c = 0;
        std::vector<double> force_vec, force_mag, extension, speed, spring_force, damping_force, react_b1, react_b2;
    };
    struct JointBuf {
        std::string name;
        chrono::ChLinkLockPrismatic* link = nullptr;
        std::vector<double> f1, t1, f2, t2;
    };
    std::vector<TsdaBuf> tsdas;
    std::vector<JointBuf> joints;
    std::vector<std::string> joint_names, tsda_names, rsda_names;
    static std::string Sanitize(const std::string& in) {     // reference :162-175
        std::string out;
        for (char ch : in) {
            if (ch == ' ') out.push_back('_');
            else if (ch == '/' || ch == '\\' || ch == ':') continue;
            else out.push_back(ch);
        }
        return out.empty() ? std::string("unnamed") : out;
    }
    void names(const std::string& p, const std::vector<std::string>& v) {
        std::vector<const char*> c;
        for (const auto& x : v) c.push_back(x.c_str());
        hc_throw_on_error(hc_h5_writer_put_string_array(w, p.c_str(), int(c.size()), c.data()));
    }

    void group(const std::string& p) { hc_throw_on_error(hc_h5_writer_put_group(w, p.c_str())); }
    void attr(const std::string& p, const std::string& n, const std::string& v) {
        hc_throw_on_error(hc_h5_writer_attr_string(w, p.c_str(), n.c_str(), v.c_str()));
    }
    void attr(const std::string& p, const std::string& n, double v) {
        hc_throw_on_error(hc_h5_writer_attr_f64(w, p.c_str(), n.c_str(), v));
    }
    void vec(const std::string& p, const std::vector<double>& v) {
        const uint64_t d[1] = {v.size()};
        hc_throw_on_error(hc_h5_writer_put_f64(w, p.c_str(), 1, d, v.data()));
    }
    void mat(const std::string& p, const std::vector<double>& v, uint64_t cols) {
        const uint64_t d[2] = {cols ? v.size() / cols : 0, cols};
        hc_throw_on_error(hc_h5_writer_put_f64(w, p.c_str(), 2, d, v.data()));
    }
    void text(const std::string& p, const std::string& s) { hc_throw_on_error(hc_h5_writer_put_string(w, p.c_str(), s.c_str())); }
};

SimulationExporter::SimulationExporter(const Options& opts) : impl_(new Impl) {
    impl_->options = opts;
    hc_throw_on_error(hc_h5_writer_create(&impl_->w));
    for (const char* g : {"/inputs", "/inputs/model", "/inputs/model/joints", "/inputs/model/tsdas", "/inputs/model/rsdas",
                          "/inputs/simulation", "/inputs/simulation/time", "/inputs/simulation/environment",
                          "/inputs/simulation/waves", "/inputs/simulation/waves/irregular", "/results", "/results/model",
                          "/results/model/bodies", "/results/model/tsdas", "/results/model/rsdas", "/results/model/joints",
                          "/results/time", "/meta", "/meta/system", "/meta/run"})
        impl_->group(g);
}

SimulationExporter::~SimulationExporter() noexcept {
    try { if (!impl_->finalized) Finalize(); } catch (...) {}
    if (impl_->w) hc_h5_writer_destroy(impl_->w);
}

void SimulationExporter::WriteSimulationInfo(chrono::ChSystem* system, const std::string& chrono_version,
                                             const std::string& model_name, double timestep, double duration_seconds) {
    if (system == nullptr) throw std::invalid_argument("WriteSimulationInfo: system must not be null");
    Impl& I = *impl_;
    I.attr("/meta", "schema_version", std::string("0.3"));
    I.attr("/meta", "files_output", I.options.output_path);
    if (!I.options.input_model_file.empty()) I.attr("/meta", "files_model", I.options.input_model_file);
    if (!I.options.input_simulation_file.empty()) I.attr("/meta", "files_simulation", I.options.input_simulation_file);
    if (!I.options.input_hydro_file.empty()) I.attr("/meta", "files_hydro", I.options.input_hydro_file);
    if (!I.options.output_tag.empty()) I.attr("/meta", "run_tag", I.options.output_tag);
    I.attr("/meta", "build_version", std::string(hc_version()));
    I.attr("/meta", "chrono_version", chrono_version);
    I.attr("/meta", "model_name", model_name);
    if (!I.options.setup_yaml_text.empty()) {
        I.text("/meta/config/setup_yaml", I.options.setup_yaml_text);
        I.attr("/meta/config", "content_type", std::string("text/yaml"));
        I.attr("/meta/config", "encoding", std::string("utf-8"));
        I.attr("/meta/config", "bytes", static_cast<double>(I.options.setup_yaml_text.size()));
    }
    I.attr("/inputs/simulation/time", "dt", timestep);
    I.attr("/inputs/simulation/time", "duration", duration_seconds);
    const auto g = system->GetGravitationalAcceleration();
    I.vec("/inputs/simulation/environment/gravity", {g.x(), g.y(), g.z()});
    I.attr("/inputs/simulation/environment", "units", std::string("m/s^2"));
    I.attr("/inputs/simulation/environment", "frame", std::string("world"));
    const std::string type = I.options.scenario_type.empty() ? std::string("still") : I.options.scenario_type;
    I.attr("/inputs/simulation/waves", "type", type);
    if (type == "regular") {
        I.attr("/inputs/simulation/waves", "H", I.options.scenario_H);
        I.attr("/inputs/simulation/waves", "T", I.options.scenario_T);
    } else if (type == "irregular") {
        I.attr("/inputs/simulation/waves", "Hs", I.options.scenario_Hs);
        I.attr("/inputs/simulation/waves", "Tp", I.options.scenario_Tp);
        if (I.options.scenario_seed >= 0) I.attr("/inputs/simulation/waves", "seed", static_cast<double>(I.options.scenario_seed));
    }
}

void SimulationExporter::WriteModel(chrono::ChSystem* system) {
    if (system == nullptr) throw std::invalid_argument("WriteModel: system must not be null");
    Impl& I = *impl_;
    I.bodies.clear();
    for (auto& b : system->GetBodies()) {
        std::string name = b->GetName();
        if (name.empty()) name = "body";
        const std::string g = "/inputs/model/bodies/" + name;
        I.group(g);
        I.attr(g, "mass", b->GetMass());
        I.attr(g, "fixed", b->IsFixed() ? 1.0 : 0.0);
        const auto p = b->GetPos();
        I.vec(g + "/location", {p.x(), p.y(), p.z()});
        const auto J = b->GetInertiaXX();
        I.vec(g + "/inertia_moments", {J.x(), J.y(), J.z()});
        I.vec(g + "/inertia_products", {0.0, 0.0, 0.0});
        I.vec(g + "/com_location", {0.0, 0.0, 0.0});
        I.vec(g + "/com_orientation", {0.0, 0.0, 0.0});
        const auto e = b->GetRot().GetCardanAnglesXYZ();
        I.vec(g + "/orientation_xyz_initial", {e.x(), e.y(), e.z()});
        I.attr(g, "orientation_xyz_initial_convention", std::string("TaitBryan_extrinsic_XYZ"));
        I.attr(g, "orientation_xyz_initial_units", std::string("rad"));
        I.text(g + "/visualization_file", "");
        Impl::BodyBuf buf;
        buf.name = name;
        I.bodies.push_back(std::move(buf));
    }
    // joints and dampers (reference :437-641)
    I.tsdas.clear(); I.joints.clear(); I.joint_names.clear(); I.tsda_names.clear(); I.rsda_names.clear();
    int tsda_idx = 0, joint_idx = 0;
    auto body_name = [](chrono::ChBody* b) { return b ? b->GetName() : std::string(); };
    for (auto& link_ptr : system->GetLinks()) {
        chrono::ChLinkBase* base = link_ptr.get();
        if (auto* tsda = dynamic_cast<chrono::ChLinkTSDA*>(base)) {
            std::string raw = base->GetName();
            if (raw.empty()) raw = "TSDA_" + std::to_string(++tsda_idx);
            const std::string nm = Impl::Sanitize(raw), g = "/inputs/model/tsdas/" + nm;
            I.tsda_names.push_back(nm);
            I.group(g);
            I.attr(g, "type", std::string("TSDA"));
            I.attr(g, "body1", body_name(tsda->GetBody1()));
            I.attr(g, "body2", body_name(tsda->GetBody2()));
            const auto p1 = tsda->GetBody1()->GetPos(), p2 = tsda->GetBody2()->GetPos();
            I.vec(g + "/point1", {p1.x(), p1.y(), p1.z()});
            I.vec(g + "/point2", {p2.x(), p2.y(), p2.z()});
            I.attr(g, "frame", std::string("world"));
            I.attr(g, "spring_coefficient", tsda->GetSpringCoefficient());
            I.attr(g, "damping_coefficient", tsda->GetDampingCoefficient());
            I.attr(g, "free_length", tsda->GetRestLength());
            Impl::TsdaBuf buf;
            buf.name = nm; buf.link = tsda; buf.rest_length = tsda->GetRestLength();
            buf.k = tsda->GetSpringCoefficient(); buf.c = tsda->GetDampingCoefficient();
            I.tsdas.push_back(std::move(buf));
            continue;
        }
        if (auto* lock = dynamic_cast<chrono::ChLinkLockPrismatic*>(base)) {
            std::string raw = base->GetName();
            if (raw.empty()) raw = "joint_" + std::to_string(++joint_idx);
            const std::string nm = Impl::Sanitize(raw), g = "/inputs/model/joints/" + nm;
            I.joint_names.push_back(nm);
            I.group(g);
            I.attr(g, "type", std::string("LOCK"));
            I.attr(g, "body1", body_name(lock->GetBody1()));
            I.attr(g, "body2", body_name(lock->GetBody2()));
            const auto loc = lock->GetBody1()->GetPos();
            const auto ax = lock->GetBody2()->GetRot().Rotate(lock->axis_in_parent);
            I.vec(g + "/location", {loc.x(), loc.y(), loc.z()});
            I.vec(g + "/axis", {ax.x(), ax.y(), ax.z()});
            I.attr(g, "frame", std::string("world"));
            Impl::JointBuf buf;
            buf.name = nm; buf.link = lock;
            I.joints.push_back(std::move(buf));
        }
    }
    // names arrays are always written, even when empty (reference :638-641)
    I.names("/inputs/model/joints/names", I.joint_names);
    I.names("/inputs/model/tsdas/names", I.tsda_names);
    I.names("/inputs/model/rsdas/names", I.rsda_names);
}

void SimulationExporter::BeginResults(chrono::ChSystem* system, int expected_steps) {
    if (system == nullptr) throw std::invalid_argument("BeginResults: system must not be null");
    if (impl_->bodies.empty()) WriteModel(system);
    impl_->time.reserve(expected_steps > 0 ? expected_steps : 0);
}

void SimulationExporter::RecordStep(chrono::ChSystem* system) {
    if (system == nullptr) throw std::invalid_argument("RecordStep: system must not be null");
    Impl& I = *impl_;
    I.time.push_back(system->GetChTime());
    const auto& cb = system->GetBodies();
    for (size_t i = 0; i < cb.size() && i < I.bodies.size(); ++i) {
        const auto& b = *cb[i];
        auto& buf = I.bodies[i];
        const auto p = b.GetPos(), v = b.GetPosDt(), a = b.GetPosDt2(), w = b.GetAngVelParent();
        const auto q = b.GetRot();
        const auto e = q.GetCardanAnglesXYZ();
        buf.pos.insert(buf.pos.end(), {p.x(), p.y(), p.z()});
        buf.vel.insert(buf.vel.end(), {v.x(), v.y(), v.z()});
        buf.acc.insert(buf.acc.end(), {a.x(), a.y(), a.z()});
        buf.quat.insert(buf.quat.end(), {q.e0, q.e1, q.e2, q.e3});
        buf.euler.insert(buf.euler.end(), {e.x(), e.y(), e.z()});
        buf.wvel.insert(buf.wvel.end(), {w.x(), w.y(), w.z()});
    }
    // TSDA channels (reference :760-794): direction body1 -> body2, speed = relative velocity along it
    for (auto& t : I.tsdas) {
        chrono::ChLinkTSDA* L = t.link;
        const double fmag = L->GetForce(), ext = L->GetLength() - t.rest_length;
        const auto p1 = L->GetBody1()->GetPos(), p2 = L->GetBody2()->GetPos();
        const auto d12 = p2 - p1;
        const double nrm = d12.Length();
        const chrono::ChVector3d dir = nrm > 1e-12 ? d12 * (1.0 / nrm) : chrono::ChVector3d(1, 0, 0);
        const auto fvec = dir * fmag;
        const double rel_speed = (L->GetBody2()->GetPosDt() - L->GetBody1()->GetPosDt()).Dot(dir);
        t.force_vec.insert(t.force_vec.end(), {fvec.x(), fvec.y(), fvec.z()});
        t.force_mag.push_back(fmag);
        t.extension.push_back(ext);
        t.speed.push_back(rel_speed);
        t.spring_force.push_back(t.k * ext);
        t.damping_force.push_back(t.c * rel_speed);
        t.react_b1.insert(t.react_b1.end(), {fvec.x(), fvec.y(), fvec.z()});
        t.react_b2.insert(t.react_b2.end(), {-fvec.x(), -fvec.y(), -fvec.z()});
    }
    // joint reactions (reference :823-851), world frame: body 2 gets the opposite force and the torque moved to it
    for (auto& j : I.joints) {
        const auto f = j.link->GetReactForce1(), tq = j.link->GetReactTorque1();
        const auto arm = j.link->GetBody1()->GetPos() - j.link->GetBody2()->GetPos();
        const auto t2 = -(tq + (arm % f));
        j.f1.insert(j.f1.end(), {f.x(), f.y(), f.z()});
        j.t1.insert(j.t1.end(), {tq.x(), tq.y(), tq.z()});
        j.f2.insert(j.f2.end(), {-f.x(), -f.y(), -f.z()});
        j.t2.insert(j.t2.end(), {t2.x(), t2.y(), t2.z()});
    }
}

void SimulationExporter::WriteIrregularInputs(const std::vector<double>& frequencies_hz, const std::vector<double>& spectral_densities,
                                              const std::vector<double>& free_surface_time, const std::vector<double>& free_surface_eta) {
    Impl& I = *impl_;
    const std::string g = "/inputs/simulation/waves/irregular";
    if (!frequencies_hz.empty()) { I.vec(g + "/frequencies_hz", frequencies_hz); I.attr(g, "frequencies_hz.units", std::string("Hz")); }
    if (!spectral_densities.empty()) {
        I.vec(g + "/spectral_densities", spectral_densities);
        I.attr(g, "spectral_densities.units", std::string("m^2/Hz"));
        I.attr(g, "spectral_densities.convention", std::string("JONSWAP (if gamma>1), else PM"));
    }
    if (!free_surface_time.empty()) { I.vec(g + "/free_surface_time", free_surface_time); I.attr(g, "free_surface_time.units", std::string("s")); }
    if (!free_surface_eta.empty()) {
        I.vec(g + "/free_surface_eta", free_surface_eta);
        I.attr(g, "free_surface_eta.units", std::string("m"));
        I.attr(g, "free_surface_eta.location", std::string("x=0,y=0,z=0 (assumed)"));
    }
}

void SimulationExporter::SetRunMetadata(const std::string& started_at_utc, const std::string& finished_at_utc, double wall_time_s,
                                        int steps, double dt_s, double time_final_s) {
    Options& o = impl_->options;
    o.run_started_at_utc = started_at_utc; o.run_finished_at_utc = finished_at_utc; o.run_wall_time_s = wall_time_s;
    o.run_steps = steps; o.run_dt = dt_s; o.run_time_final = time_final_s;
}

void SimulationExporter::Finalize() {
    Impl& I = *impl_;
    if (I.finalized) return;
    I.vec("/results/time/time", I.time);
    I.attr("/results/time", "units", std::string("s"));
    for (auto& b : I.bodies) {
        const std::string g = "/results/model/bodies/" + b.name;
        I.group(g);
        I.mat(g + "/position", b.pos, 3);
        I.mat(g + "/velocity", b.vel, 3);
        I.mat(g + "/acceleration", b.acc, 3);
        I.mat(g + "/orientation", b.quat, 4);
        I.mat(g + "/orientation_xyz", b.euler, 3);
        I.mat(g + "/angular_velocity", b.wvel, 3);
        I.attr(g, "position_units", std::string("m"));
        I.attr(g, "position_frame", std::string("world"));
        I.attr(g, "velocity_units", std::string("m/s"));
        I.attr(g, "acceleration_units", std::string("m/s^2"));
        I.attr(g, "orientation_order", std::string("wxyz"));
        I.attr(g, "orientation_xyz_convention", std::string("TaitBryan_extrinsic_XYZ"));
        I.attr(g, "orientation_xyz_units", std::string("rad"));
        I.attr(g, "angular_velocity_units", std::string("rad/s"));
    }
    for (auto& t : I.tsdas) {
        const std::string g = "/results/model/tsdas/" + t.name;
        I.group(g);
        I.attr(g, "type", std::string("TSDA"));
        I.attr(g, "time_ref", std::string("/results/time/time"));
        I.attr(g, "frame", std::string("world"));
        I.attr(g, "units_force", std::string("N"));
        I.attr(g, "units_extension", std::string("m"));
        I.attr(g, "units_speed", std::string("m/s"));
        I.mat(g + "/force_vec", t.force_vec, 3);
        I.vec(g + "/force_mag", t.force_mag);
        I.vec(g + "/extension", t.extension);
        I.vec(g + "/speed", t.speed);
        I.vec(g + "/spring_force", t.spring_force);
        I.vec(g + "/damping_force", t.damping_force);
        I.mat(g + "/reaction_force_body1", t.react_b1, 3);
        I.mat(g + "/reaction_force_body2", t.react_b2, 3);
    }
    for (auto& j : I.joints) {
        const std::string g = "/results/model/joints/" + j.name;
        I.group(g);
        I.attr(g, "type", std::string("LOCK"));
        I.attr(g, "class", std::string("ChLinkLockPrismatic"));
        I.attr(g, "time_ref", std::string("/results/time/time"));
        I.attr(g, "frame1", std::string("world"));
        I.attr(g, "frame2", std::string("world"));
        I.attr(g, "units_force", std::string("N"));
        I.attr(g, "units_torque", std::string("N*m"));
        I.mat(g + "/reaction1_force", j.f1, 3);
        I.mat(g + "/reaction1_torque", j.t1, 3);
        I.mat(g + "/reaction2_force", j.f2, 3);
        I.mat(g + "/reaction2_torque", j.t2, 3);
    }
    const Options& o = I.options;
    if (!o.run_started_at_utc.empty()) I.attr("/meta/run", "started_at_utc", o.run_started_at_utc);
    if (!o.run_finished_at_utc.empty()) I.attr("/meta/run", "finished_at_utc", o.run_finished_at_utc);
    I.attr("/meta/run", "wall_time_s", o.run_wall_time_s);
    I.attr("/meta/run", "steps", static_cast<double>(o.run_steps ? o.run_steps : int(I.time.size())));
    I.attr("/meta/run", "dt_s", o.run_dt);
    I.attr("/meta/run", "time_final_s", o.run_time_final);
    hc_throw_on_error(hc_h5_writer_save(I.w, o.output_path.c_str()));
    I.finalized = true;
}

}  // namespace hydroc
