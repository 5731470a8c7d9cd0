// Host helpers (reference src/helper.cpp).
#include <hydroc/helper.h>

#include <algorithm>
#include <cstdlib>
#include <filesystem>
#include <iostream>
#include <stdexcept>

size_t get_lower_index(double value, const std::vector<double>& ticks) {
    // index i with ticks[i] < value <= ticks[i+1]; the first and the last interval are rejected
    size_t idx = size_t(std::upper_bound(ticks.begin(), ticks.end(), value) - ticks.begin()) - 1;
    if (idx < ticks.size() && ticks[idx] == value) idx -= 1;
    if (idx == 0 || idx >= ticks.size() - 1)
        throw std::runtime_error("Could not find index for value " + std::to_string(value) + " in array with bounds (" +
                                 std::to_string(ticks.front()) + ", " + std::to_string(ticks.back()) + ").");
    return idx;
}

namespace hydroc {
static std::filesystem::path g_datadir;

int SetInitialEnvironment(int argc, char* argv[]) noexcept {
    try {
        const char* env = std::getenv("HYDROCHRONO_DATA_DIR");
        if (env) g_datadir = std::filesystem::absolute(env);
        else if (argc >= 2) g_datadir = std::filesystem::absolute(argv[1]);
        else {
            std::cerr << "Usage: .exe [<datadir>] or set HYDROCHRONO_DATA_DIR environment variable" << std::endl;
            g_datadir = std::filesystem::absolute(std::filesystem::path("..") / ".." / "demos");
        }
    } catch (...) {
        return 1;
    }
    return 0;
}
std::string getDataDir() noexcept { return g_datadir.lexically_normal().generic_string(); }
}  // namespace hydroc
