// maps C-ABI status codes back to the exception types the reference throws
#pragma once
#include <stdexcept>
#include <string>

#include <hydrochrono_b200.h>

inline void hc_throw_on_error(hc_status s) {
    if (s == HC_OK) return;
    const std::string msg = hc_last_error();
    if (s == HC_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);
    throw std::runtime_error(msg);
}
