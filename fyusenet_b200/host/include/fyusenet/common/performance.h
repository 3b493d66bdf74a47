// Monotonic-clock helpers (reference: fyusenet/common/performance.cpp:29-49).
#pragma once
#include <chrono>
#include <cstdint>

using tstamp = std::chrono::steady_clock::time_point;
inline tstamp fy_get_stamp() { return std::chrono::steady_clock::now(); }
inline uint32_t fy_elapsed_millis(const tstamp &a, const tstamp &b) {
    return (uint32_t)std::chrono::duration_cast<std::chrono::milliseconds>(b - a).count();
}
inline uint64_t fy_elapsed_micros(const tstamp &a, const tstamp &b) {
    return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count();
}
