// YAML -> WaveBase + TestHydro (reference src/setup_hydro_from_yaml.h:33-39).
#ifndef HYDROC_B200_SETUP_HYDRO_FROM_YAML_H
#define HYDROC_B200_SETUP_HYDRO_FROM_YAML_H

#include <memory>
#include <vector>

#include <chrono_compat/chrono_compat.h>
#include <hydroc/hydro_types.h>

class TestHydro;

std::unique_ptr<TestHydro> SetupHydroFromYAML(const YAMLHydroData& hydro_data,
                                              const std::vector<std::shared_ptr<chrono::ChBody>>& bodies,
                                              double timestep, double sim_duration, double ramp_duration);

#endif
