// H5FileInfo / HydroData over the C-ABI tables handle (reference src/h5fileinfo.cpp).
#include <hydroc/h5fileinfo.h>

#include <cmath>
#include <filesystem>
#include <iostream>

#include "hc_check.h"

H5FileInfo::H5FileInfo(std::string file, int num_bod) : h5_file_name_(std::move(file)), num_bodies_(num_bod) {
    if (!std::filesystem::exists(h5_file_name_))   // src/h5fileinfo.cpp:20-24 (warning only; reading throws)
        std::cerr << "WARNING: H5 file does not exist, absolute file location: "
                  << std::filesystem::absolute(h5_file_name_).string() << std::endl;
}
H5FileInfo::~H5FileInfo() {}

HydroData H5FileInfo::ReadH5Data() {
    hc_tables* t = nullptr;
    hc_throw_on_error(hc_tables_load_h5(h5_file_name_.c_str(), num_bodies_, &t));
    return FromTables(t, h5_file_name_);
}

HydroData H5FileInfo::FromTables(hc_tables* t, const std::string& label) {
    HydroData d;
    d.tables_ = std::shared_ptr<hc_tables>(t, hc_tables_destroy);
    const int N = hc_tables_num_bodies(t), D = 6 * N, L = hc_tables_rirf_steps(t);
    const int nw = hc_tables_num_freqs(t), Le0 = hc_tables_exc_irf_steps(t);
    d.sim_data_.h5_file_name = label;
    d.sim_data_.rho = hc_tables_rho(t);
    d.sim_data_.g = hc_tables_g(t);
    d.sim_data_.water_depth = hc_tables_water_depth(t);
    d.body_data_.resize(N);
    d.reg_wave_data_.resize(N);
    d.irreg_wave_data_.resize(N);
    for (int b = 0; b < N; ++b) {
        HydroData::BodyInfo& B = d.body_data_[b];
        B.body_name = "body" + std::to_string(b + 1);
        B.body_num = b;
        hc_throw_on_error(hc_tables_disp_vol(t, b, &B.disp_vol));
        B.rirf_time_vector.resize(L);
        hc_throw_on_error(hc_tables_rirf_time(t, B.rirf_time_vector.data()));
        B.rirf_timestep = B.rirf_time_vector[1] - B.rirf_time_vector[0];
        B.cg.resize(3); B.cb.resize(3);
        hc_throw_on_error(hc_tables_cg(t, b, B.cg.data()));
        hc_throw_on_error(hc_tables_cb(t, b, B.cb.data()));
        B.lin_matrix.resize(6, 6);
        hc_throw_on_error(hc_tables_lin_matrix(t, b, B.lin_matrix.data()));
        B.inf_added_mass.resize(6, D);
        hc_throw_on_error(hc_tables_inf_added_mass(t, b, B.inf_added_mass.data()));
        if (nw > 0) {
            HydroData::RegularWaveInfo& R = d.reg_wave_data_[b];
            R.freq_list.resize(nw);
            hc_throw_on_error(hc_tables_freq_list(t, R.freq_list.data()));
            R.excitation_mag_matrix.resize(6, nw);
            R.excitation_phase_matrix.resize(6, nw);
            hc_throw_on_error(hc_tables_excitation_mag(t, b, R.excitation_mag_matrix.data()));
            hc_throw_on_error(hc_tables_excitation_phase(t, b, R.excitation_phase_matrix.data()));
        }
        if (Le0 > 0) {
            HydroData::IrregularWaveInfo& I = d.irreg_wave_data_[b];
            I.excitation_irf_time.resize(Le0);
            I.excitation_irf_matrix.resize(6, Le0);
            hc_throw_on_error(hc_tables_excitation_irf(t, b, I.excitation_irf_time.data(), I.excitation_irf_matrix.data()));
        }
    }
    return d;
}

Eigen::MatrixXd HydroData::GetInfAddedMassMatrix(int b) const { return body_data_.at(b).inf_added_mass; }

double HydroData::GetHydrostaticStiffnessVal(int b, int i, int j) const {
    double v = 0;
    hc_throw_on_error(hc_tables_hydrostatic_stiffness(tables_.get(), b, i, j, &v));
    return v;
}

Eigen::MatrixXd HydroData::GetLinMatrix(int b) const { return body_data_.at(b).lin_matrix; }

double HydroData::GetRIRFVal(int b, int dof, int col, int s) const {
    double v = 0;
    hc_throw_on_error(hc_tables_rirf_val(tables_.get(), 6 * b + dof, col, s, &v));
    return v;
}

int HydroData::GetRIRFDims(int i) const {
    if (i == 0) return 6;
    if (i == 1) return 6 * int(body_data_.size());
    return hc_tables_rirf_steps(tables_.get());
}

Eigen::VectorXd HydroData::GetRIRFTimeVector() const { return body_data_.at(0).rirf_time_vector; }
