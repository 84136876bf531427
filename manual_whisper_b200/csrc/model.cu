// mw_model lifetime and the encoder driver (S2: ctranslate2.models.Whisper.encode).
//
// Encoder data flow, one launch list per call (B chunks, T = n_audio_ctx, F = 2T frames):
//   features (bf16 time-major [B, F+2, n_mels], zero edge rows)
//   conv1  : implicit GEMM, A row t = 3*n_mels contiguous elements starting at padded row t    -> h1 (+bias, GELU)
//   conv2  : implicit GEMM, stride 2: A row t = 3*d contiguous elements starting at padded row 2t -> x = GELU(.)+pos (fp32)
//   N x { LN -> QKV GEMM -> flash attention -> out GEMM (+residual) -> LN -> fc1 GEMM (GELU) -> fc2 GEMM (+residual) }
//   final LN -> bf16 [B, T, d]
#include "model.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace mw {

mw_status model_alloc(mw_model* m, void** ptr, int64_t bytes, bool zero) {
    MW_CUDA_CHECK(cudaMalloc(ptr, (size_t)bytes));
    m->allocations.push_back(*ptr);
    m->workspace_bytes += bytes;
    if (zero) MW_CUDA_CHECK(cudaMemset(*ptr, 0, (size_t)bytes));
    return MW_OK;
}

static mw_status encode_from_time_major(mw_model* m, const mw_h* mel_t, int B, void* d_enc_out, cudaStream_t st) {
    const mw_model_config& c = m->cfg;
    const int T = c.n_audio_ctx, F = 2 * T, d = c.d_model;
    mw_status s;
    {   // conv1 + GELU -> h1 rows 1..F
        GemmArgs a;
        a.a = mel_t; a.a_row_stride = c.n_mels; a.a_batch_stride = (int64_t)(F + 2) * c.n_mels;
        a.w = m->gw(MW_W_CONV1); a.w_row_stride = 3 * c.n_mels;
        a.bias = (const float*)m->gw(MW_B_CONV1);
        a.out = m->h1; a.out_batch_rows = F + 2; a.out_row_off = 1; a.ld_out = d;
        a.batch = B; a.M = F; a.N = d; a.K = 3 * c.n_mels; a.gelu = true; a.out_f32 = false;
        if ((s = gemm_launch(a, st)) != MW_OK) return s;
    }
    {   // conv2 (stride 2) + GELU + positional embedding -> x (fp32)
        GemmArgs a;
        a.a = m->h1; a.a_row_stride = 2 * d; a.a_batch_stride = (int64_t)(F + 2) * d;
        a.w = m->gw(MW_W_CONV2); a.w_row_stride = 3 * d;
        a.bias = (const float*)m->gw(MW_B_CONV2);
        a.residual = (const float*)m->gw(MW_ENC_POS); a.res_batch_rows = 0; a.ld_res = d;
        a.out = m->x; a.out_batch_rows = T; a.out_row_off = 0; a.ld_out = d;
        a.batch = B; a.M = T; a.N = d; a.K = 3 * d; a.gelu = true; a.out_f32 = true;
        if ((s = gemm_launch(a, st)) != MW_OK) return s;
    }
    const int rows = B * T;
    for (int l = 0; l < c.enc_layers; ++l) {
        if ((s = layernorm_launch(m->x, (const float*)m->elw(l, MW_EL_LN1_G), (const float*)m->elw(l, MW_EL_LN1_B), m->ln, rows, d, st)) != MW_OK) return s;
        {
            GemmArgs a;
            a.a = m->ln; a.a_row_stride = d; a.w = m->elw(l, MW_EL_WQKV); a.w_row_stride = d;
            a.bias = (const float*)m->elw(l, MW_EL_BQKV);
            a.out = m->qkv; a.ld_out = 3 * d; a.M = rows; a.N = 3 * d; a.K = d;
            if ((s = gemm_launch(a, st)) != MW_OK) return s;
        }
        if ((s = attention_launch(m->qkv, m->att, B, T, c.n_heads, st)) != MW_OK) return s;
        {
            GemmArgs a;
            a.a = m->att; a.a_row_stride = d; a.w = m->elw(l, MW_EL_WO); a.w_row_stride = d;
            a.bias = (const float*)m->elw(l, MW_EL_BO);
            a.residual = m->x; a.ld_res = d;
            a.out = m->x; a.ld_out = d; a.M = rows; a.N = d; a.K = d; a.out_f32 = true;
            if ((s = gemm_launch(a, st)) != MW_OK) return s;
        }
        if ((s = layernorm_launch(m->x, (const float*)m->elw(l, MW_EL_LN2_G), (const float*)m->elw(l, MW_EL_LN2_B), m->ln, rows, d, st)) != MW_OK) return s;
        {
            GemmArgs a;
            a.a = m->ln; a.a_row_stride = d; a.w = m->elw(l, MW_EL_W1); a.w_row_stride = d;
            a.bias = (const float*)m->elw(l, MW_EL_B1);
            a.out = m->mlp; a.ld_out = c.ffn; a.M = rows; a.N = c.ffn; a.K = d; a.gelu = true;
            if ((s = gemm_launch(a, st)) != MW_OK) return s;
        }
        {
            GemmArgs a;
            a.a = m->mlp; a.a_row_stride = c.ffn; a.w = m->elw(l, MW_EL_W2); a.w_row_stride = c.ffn;
            a.bias = (const float*)m->elw(l, MW_EL_B2);
            a.residual = m->x; a.ld_res = d;
            a.out = m->x; a.ld_out = d; a.M = rows; a.N = d; a.K = c.ffn; a.out_f32 = true;
            if ((s = gemm_launch(a, st)) != MW_OK) return s;
        }
    }
    return layernorm_launch(m->x, (const float*)m->gw(MW_ENC_LN_G), (const float*)m->gw(MW_ENC_LN_B), d_enc_out, rows, d, st);
}

}  // namespace mw

extern "C" mw_status mw_model_create(const mw_model_config* cfg, const mw_weight_table* weights, mw_model** out_model) {
    MW_REQUIRE(cfg && weights && out_model, "mw_model_create: null argument");
    MW_REQUIRE(cfg->d_model == cfg->n_heads * 64, "mw_model_create: d_head must be 64 (d_model=%d n_heads=%d)", cfg->d_model, cfg->n_heads);
    MW_REQUIRE(cfg->d_model % 128 == 0 && cfg->ffn % 128 == 0, "mw_model_create: d_model and ffn must be multiples of 128");
    MW_REQUIRE(cfg->n_mels % 8 == 0 && cfg->n_mels > 0, "mw_model_create: n_mels must be a positive multiple of 8");
    MW_REQUIRE(cfg->max_batch > 0 && cfg->max_beam > 0 && cfg->max_beam <= 8, "mw_model_create: bad max_batch/max_beam");
    MW_REQUIRE(cfg->n_audio_ctx > 0 && cfg->n_text_ctx > 1 && cfg->vocab > 0, "mw_model_create: bad context sizes");
    const int expect = MW_GLOBAL_COUNT + cfg->enc_layers * MW_EL_COUNT + cfg->dec_layers * MW_DL_COUNT;
    MW_REQUIRE(weights->n == expect, "mw_model_create: weight table has %d entries, expected %d", weights->n, expect);
    for (int i = 0; i < expect; ++i) MW_REQUIRE(weights->ptrs[i] != nullptr, "mw_model_create: weight %d is null", i);
    mw::DeviceGuard guard(cfg->device);
    mw_model* m = new mw_model();
    m->cfg = *cfg;
    m->w.assign(weights->ptrs, weights->ptrs + expect);
    const int64_t B = cfg->max_batch, T = cfg->n_audio_ctx, F = 2 * T, d = cfg->d_model;
    mw_status s = MW_OK;
    auto A = [&](void** p, int64_t bytes, bool zero) { if (s == MW_OK) s = mw::model_alloc(m, p, bytes, zero); };
    A((void**)&m->mel_t, B * (F + 2) * cfg->n_mels * 2, true);
    A((void**)&m->h1, B * (F + 2) * d * 2, true);
    A((void**)&m->x, B * T * d * 4, false);
    A((void**)&m->ln, B * T * d * 2, false);
    A((void**)&m->qkv, B * T * 3 * d * 2, false);
    A((void**)&m->att, B * T * d * 2, false);
    A((void**)&m->mlp, B * T * cfg->ffn * 2, false);
    if (s == MW_OK) s = mw::decoder_state_create(m);
    if (s != MW_OK) { mw_model_destroy(m); return s; }
    *out_model = m;
    return MW_OK;
}

extern "C" void mw_model_destroy(mw_model* m) {
    if (!m) return;
    mw::DeviceGuard guard(m->cfg.device);
    cudaDeviceSynchronize();
    mw::decoder_state_destroy(m);
    for (void* p : m->allocations) cudaFree(p);
    delete m;
}

extern "C" int64_t mw_model_workspace_bytes(const mw_model* m) { return m ? m->workspace_bytes : 0; }

extern "C" mw_status mw_encode_t(mw_model* m, const void* d_mel_t, int B, void* d_enc_out, void* stream) {
    MW_REQUIRE(m && d_mel_t && d_enc_out, "mw_encode_t: null argument");
    MW_REQUIRE(B > 0 && B <= m->cfg.max_batch, "mw_encode_t: B=%d outside 1..max_batch=%d", B, m->cfg.max_batch);
    mw::DeviceGuard guard(m->cfg.device);
    return mw::encode_from_time_major(m, (const mw_h*)d_mel_t, B, d_enc_out, (cudaStream_t)stream);
}

extern "C" mw_status mw_encode(mw_model* m, const float* d_mel, int B, void* d_enc_out, void* stream) {
    MW_REQUIRE(m && d_mel && d_enc_out, "mw_encode: null argument");
    MW_REQUIRE(B > 0 && B <= m->cfg.max_batch, "mw_encode: B=%d outside 1..max_batch=%d", B, m->cfg.max_batch);
    mw::DeviceGuard guard(m->cfg.device);
    mw_status s = mw::features_to_time_major_launch(d_mel, m->mel_t, B, m->cfg.n_mels, m->frames(), (cudaStream_t)stream);
    if (s != MW_OK) return s;
    return mw::encode_from_time_major(m, m->mel_t, B, d_enc_out, (cudaStream_t)stream);
}
