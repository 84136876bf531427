// Entry points of include/mw_b200.h that are not implemented yet return MW_ERR_UNSUPPORTED (never a
// silent fallback).  This file shrinks to nothing as the engine lands.
#include "mw_common.cuh"

#define MW_STUB(name) mw::set_error(#name ": not implemented in this build"); return MW_ERR_UNSUPPORTED

extern "C" mw_status mw_model_create(const mw_model_config*, const mw_weight_table*, mw_model**) { MW_STUB(mw_model_create); }
extern "C" void mw_model_destroy(mw_model*) {}
extern "C" int64_t mw_model_workspace_bytes(const mw_model*) { return 0; }
extern "C" mw_status mw_encode(mw_model*, const float*, int, void*, void*) { MW_STUB(mw_encode); }
extern "C" mw_status mw_encode_t(mw_model*, const void*, int, void*, void*) { MW_STUB(mw_encode_t); }
extern "C" mw_status mw_generate(mw_model*, const void*, int, const int32_t*, int, const mw_gen_options*, int32_t*,
                                 int32_t*, float*, void*) { MW_STUB(mw_generate); }
extern "C" mw_status mw_decoder_logits(mw_model*, const void*, int, const int32_t*, int, float*, void*) { MW_STUB(mw_decoder_logits); }
extern "C" mw_status mw_detect_language(mw_model*, const void*, int, int32_t, int32_t, int32_t, float*, void*) { MW_STUB(mw_detect_language); }
extern "C" mw_status mw_attention_bf16(const void*, void*, int, int, int, void*) { MW_STUB(mw_attention_bf16); }
extern "C" mw_status mw_layernorm(const float*, const float*, const float*, void*, int, int, void*) { MW_STUB(mw_layernorm); }
