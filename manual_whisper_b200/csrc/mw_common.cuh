// Shared host-side plumbing of libmw_b200.so: error strings, launch counting, device guards.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>
#include <cstdlib>
#include <mutex>

#include "../../include/mw_b200.h"

// ---- 16-bit storage type of activations, K/V caches and weight matrices.
// Default fp16: it is the reference's own GPU compute type (compute_type="float16", /root/reference/transcribe_colab.ipynb:119),
// Whisper checkpoints are published in fp16 (so weights are exact), and its 11-bit significand puts the engine's logit noise
// 8x below bf16's - which is what lets greedy ids match the fp32 oracle.  Accumulation, the residual
// stream, softmax statistics and logits are fp32 either way.  -DMW_STORAGE_BF16 builds the bf16 variant for A/B runs.
#ifdef MW_STORAGE_BF16
typedef __nv_bfloat16 mw_h;
typedef __nv_bfloat162 mw_h2;
#define MW_STORAGE_NAME "bf16"
#define MW_MMA_SYNC_TYPE "bf16"
#define MW_UMMA_FMT 1u
__device__ __forceinline__ mw_h f2h(float v) { return __float2bfloat16(v); }
__device__ __forceinline__ float h2f(mw_h v) { return __bfloat162float(v); }
__device__ __forceinline__ mw_h2 f2h2(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ float2 h22f2(mw_h2 v) { return __bfloat1622float2(v); }
__device__ __forceinline__ mw_h2 f2h2_bounded(float a, float b) { return __floats2bfloat162_rn(a, b); }
#else
typedef __half mw_h;
typedef __half2 mw_h2;
#define MW_STORAGE_NAME "fp16"
#define MW_MMA_SYNC_TYPE "f16"
#define MW_UMMA_FMT 0u
// saturating: a value beyond the fp16 range becomes +-65504, never inf (which would turn into NaN downstream)
__device__ __forceinline__ mw_h f2h(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f)); }
__device__ __forceinline__ float h2f(mw_h v) { return __half2float(v); }
__device__ __forceinline__ mw_h2 f2h2(float a, float b) {
    return __floats2half2_rn(fminf(fmaxf(a, -65504.0f), 65504.0f), fminf(fmaxf(b, -65504.0f), 65504.0f));
}
__device__ __forceinline__ float2 h22f2(mw_h2 v) { return __half22float2(v); }
// for values known to lie inside the fp16 range (softmax probabilities, convex combinations of stored values): no clamp
__device__ __forceinline__ mw_h2 f2h2_bounded(float a, float b) { return __floats2half2_rn(a, b); }
#endif

namespace mw {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

// Launches enqueued while this thread captures a CUDA graph are not launches yet: they are tallied per graph
// (t_captured) and added to g_launches each time the graph is replayed.
extern thread_local bool t_capturing;
extern thread_local uint64_t t_captured;
inline void count_launch(int n = 1) {
    if (t_capturing) t_captured += (uint64_t)n;
    else g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed);
}

// Programmatic dependent launch (the decode step is a chain of ~350 short dependent kernels): a kernel launched through
// launch_chained() may be scheduled while its predecessor in the stream is still running.  It calls pdl_trigger() first (its
// own successor may be scheduled as well) and pdl_wait() before it touches anything an earlier kernel of the chain writes -
// pdl_wait() returns when the predecessor grid has completed and flushed, which transitively covers everything before it.
// Whatever precedes pdl_wait() may only READ data that is constant during the chain (weights, encoder K/V) and write
// nothing.  Both instructions are no-ops in a kernel launched the ordinary way.
// OFF by default (MW_PDL=1 turns the attribute on): inside the step's CUDA graph it measured 3.84 vs 4.05 us per node on a
// chain of stand-alone LayerNorms, but the real single-stream step got 4 % SLOWER (4.06 vs 3.89 ms) and eight concurrent
// batches did not move - graph edges between kernel nodes are already cheap, and the early CTAs only take SM slots.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("MW_PDL"); return e && e[0] == '1'; }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        // always set: on a fresh host thread this is what binds the primary context, which the driver-API tensor-map
        // encoder needs (a thread whose first call into the library is mw_generate got CUDA_ERROR_INVALID_CONTEXT)
        cudaSetDevice(dev);
        if (prev == dev) prev = -1;
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
