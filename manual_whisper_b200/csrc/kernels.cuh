// Launchers shared between the engine's translation units.
#pragma once
#include "mw_common.cuh"

namespace mw {

mw_status layernorm_launch(const float* x, const float* gamma, const float* beta, void* out_bf16, int rows, int d,
                           cudaStream_t st);
mw_status features_to_time_major_launch(const float* in, void* out_bf16, int B, int C, int F, cudaStream_t st);
mw_status attention_launch(const void* d_qkv, void* d_out, int B, int T, int n_heads, cudaStream_t st);
mw_status attention_launch_general(const void* d_q, int64_t ldq, int colq0, const void* d_kv, int64_t ldkv, int colk0, int colv0,
                                   void* d_out, int out_ld, int B, int Tq, int Tk, int n_heads, cudaStream_t st);

}  // namespace mw
