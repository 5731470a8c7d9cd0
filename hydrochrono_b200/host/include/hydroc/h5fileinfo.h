// H5FileInfo / HydroData with the reference's public surface (include/hydroc/h5fileinfo.h:35-260), backed by
// the C-ABI tables handle (hc_tables) of libhydrochrono_b200: the BEMIO .h5 is parsed by the library's own
// classic-HDF5 reader and the tables are what gets staged into HBM.
#ifndef HYDROC_B200_H5FILEINFO_H
#define HYDROC_B200_H5FILEINFO_H
#pragma once

#include <memory>
#include <string>
#include <vector>

#include <chrono_compat/eigen_compat.h>
#include <hydrochrono_b200.h>

class H5FileInfo;

class HydroData {
  public:
    struct BodyInfo {
        std::string body_name;
        int body_num = 0;
        double disp_vol = 0;
        Eigen::VectorXd rirf_time_vector;
        double rirf_timestep = 0;
        Eigen::VectorXd cg;
        Eigen::VectorXd cb;
        Eigen::MatrixXd lin_matrix;       // 6 x 6, unscaled
        Eigen::MatrixXd inf_added_mass;   // 6 x 6N, x rho
    };
    struct SimulationParameters {
        std::string h5_file_name;
        double rho = 0;
        double g = 0;
        double water_depth = 0;
    };
    struct RegularWaveInfo {
        Eigen::VectorXd freq_list;
        Eigen::MatrixXd excitation_mag_matrix;    // 6 x nw (wave direction 0), x rho*g
        Eigen::MatrixXd excitation_phase_matrix;  // 6 x nw
    };
    struct IrregularWaveInfo {
        Eigen::VectorXd excitation_irf_time;
        Eigen::MatrixXd excitation_irf_matrix;    // 6 x Le0, x rho*g
    };

    // getters: body index first, 0-based (include/hydroc/h5fileinfo.h:102-225)
    Eigen::MatrixXd GetInfAddedMassMatrix(int b) const;
    double GetHydrostaticStiffnessVal(int b, int i, int j) const;
    Eigen::MatrixXd GetLinMatrix(int b) const;
    double GetRIRFVal(int b, int dof, int col, int s) const;
    double GetDispVolVal(int b) const { return body_data_.at(b).disp_vol; }
    Eigen::VectorXd GetCGVector(int b) const { return body_data_.at(b).cg; }
    Eigen::VectorXd GetCBVector(int b) const { return body_data_.at(b).cb; }
    int GetRIRFDims(int i) const;
    Eigen::VectorXd GetRIRFTimeVector() const;
    double GetRhoVal() const { return sim_data_.rho; }
    std::vector<BodyInfo>& GetBodyInfos() { return body_data_; }
    SimulationParameters& GetSimulationInfo() { return sim_data_; }
    std::vector<RegularWaveInfo>& GetRegularWaveInfos() { return reg_wave_data_; }
    std::vector<IrregularWaveInfo>& GetIrregularWaveInfos() { return irreg_wave_data_; }

    // the device-staging handle behind this HydroData (shared with TestHydro / ChLoadAddedMass)
    hc_tables* handle() const { return tables_.get(); }
    int num_bodies() const { return int(body_data_.size()); }

  private:
    friend class H5FileInfo;
    HydroData() = default;
    std::shared_ptr<hc_tables> tables_;
    std::vector<BodyInfo> body_data_;
    SimulationParameters sim_data_;
    std::vector<RegularWaveInfo> reg_wave_data_;
    std::vector<IrregularWaveInfo> irreg_wave_data_;
};

class H5FileInfo {
  public:
    H5FileInfo(std::string file, int num_bod = 1);
    H5FileInfo() = delete;
    ~H5FileInfo();
    // Reads the BEMIO file (throws std::runtime_error like the reference, src/h5fileinfo.cpp:172-181)
    HydroData ReadH5Data();
    // Builds a HydroData around tables created from in-memory arrays (hc_tables_create); takes ownership.
    static HydroData FromTables(hc_tables* tables, const std::string& label = "");

  private:
    std::string h5_file_name_;
    int num_bodies_;
};

#endif
