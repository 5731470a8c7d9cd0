// Helper functions of the reference kept on the host (include/hydroc/helper.h).
#ifndef HYDROC_B200_HELPER_H
#define HYDROC_B200_HELPER_H
#pragma once

#include <cstddef>
#include <string>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// nearest-below bracket index used by the excitation convolution (src/helper.cpp:8-22); throws std::runtime_error
// when the value falls into the first or beyond the last interval
size_t get_lower_index(double value, const std::vector<double>& ticks);

namespace hydroc {
// HYDROCHRONO_DATA_DIR / argv[1] resolution of the demos (src/helper.cpp:24-48)
int SetInitialEnvironment(int argc, char* argv[]) noexcept;
std::string getDataDir() noexcept;
}  // namespace hydroc

#endif
