// Shared host-side plumbing of libmw_b200.so: error strings, launch counting, device guards.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>
#include <mutex>

#include "../../include/mw_b200.h"

namespace mw {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define MW_CUDA_CHECK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            mw::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return MW_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define MW_LAUNCH_CHECK()                                                                        \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            mw::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return MW_ERR_CUDA;                                                                  \
        }                                                                                        \
        mw::count_launch();                                                                      \
    } while (0)

#define MW_REQUIRE(cond, ...)                                                                    \
    do {                                                                                         \
        if (!(cond)) { mw::set_error(__VA_ARGS__); return MW_ERR_INVALID; }                      \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute is per device and must have completed before ANY thread launches the kernel there
// (several host threads drive replicas concurrently): a mutex-guarded per-device flag.
struct PerDeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    template <typename Fn>
    cudaError_t run(Fn&& fn) {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        std::lock_guard<std::mutex> lock(mu);
        if (done[dev]) return cudaSuccess;
        cudaError_t e = fn();
        if (e == cudaSuccess) done[dev] = true;
        return e;
    }
};
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mw
