// shared by the reference-style demo mains: trajectory output in the reference's text format
#pragma once
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>

inline int write_heave(const std::string& path, const std::vector<double>& t, const std::vector<double>& z) {
    std::ofstream out(path);
    if (!out.is_open()) return 1;
    out << std::left << std::setw(10) << "Time (s)" << std::right << std::setw(12) << "Heave (m)" << std::endl;
    for (size_t i = 0; i < t.size(); ++i)
        out << std::left << std::setw(12) << std::setprecision(6) << std::fixed << t[i] << std::right << std::setw(12)
            << std::setprecision(6) << std::fixed << z[i] << std::endl;
    return 0;
}
