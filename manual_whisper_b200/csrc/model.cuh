// The opaque mw_model: borrowed weight pointers + the workspace sized once at create (SURVEY.md §8b
// "Ownership": no allocation on the hot path).
#pragma once
#include "mw_common.cuh"
#include <vector>

struct mw_model {
    mw_model_config cfg{};
    std::vector<const void*> w;          // copy of the caller's weight table (device pointers, borrowed)
    int64_t workspace_bytes = 0;
    std::vector<void*> allocations;

    // ---- encoder workspace (max_batch chunks)
    mw_h* mel_t = nullptr;      // [B, F+2, n_mels]   F = 2*n_audio_ctx
    mw_h* h1 = nullptr;         // [B, F+2, d]        conv1 output, rows 0 and F+1 stay zero
    float* x = nullptr;                  // [B*T, d]           fp32 residual stream
    mw_h* ln = nullptr;         // [B*T, d]
    mw_h* qkv = nullptr;        // [B*T, 3d]
    mw_h* att = nullptr;        // [B*T, d]
    mw_h* mlp = nullptr;        // [B*T, ffn]

    // ---- decoder state (decoder.cu)
    struct DecoderState* dec = nullptr;

    const void* gw(int id) const { return w[id]; }
    const void* elw(int layer, int id) const { return w[MW_GLOBAL_COUNT + layer * MW_EL_COUNT + id]; }
    const void* dlw(int layer, int id) const {
        return w[MW_GLOBAL_COUNT + cfg.enc_layers * MW_EL_COUNT + layer * MW_DL_COUNT + id];
    }
    int frames() const { return 2 * cfg.n_audio_ctx; }
};

namespace mw {
mw_status model_alloc(mw_model* m, void** ptr, int64_t bytes, bool zero);
mw_status decoder_state_create(mw_model* m);
void decoder_state_destroy(mw_model* m);
}  // namespace mw
