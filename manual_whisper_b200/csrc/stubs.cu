// Entry points of include/mw_b200.h that are not implemented yet return MW_ERR_UNSUPPORTED (never a
// silent fallback).  This file shrinks to nothing as the engine lands.
#include "mw_common.cuh"
#include "model.cuh"

#define MW_STUB(name) mw::set_error(#name ": not implemented in this build"); return MW_ERR_UNSUPPORTED

extern "C" mw_status mw_generate(mw_model*, const void*, int, const int32_t*, int, const mw_gen_options*, int32_t*,
                                 int32_t*, float*, void*) { MW_STUB(mw_generate); }
extern "C" mw_status mw_decoder_logits(mw_model*, const void*, int, const int32_t*, int, float*, void*) { MW_STUB(mw_decoder_logits); }
extern "C" mw_status mw_detect_language(mw_model*, const void*, int, int32_t, int32_t, int32_t, float*, void*) { MW_STUB(mw_detect_language); }

namespace mw {
mw_status decoder_state_create(mw_model*) { return MW_OK; }
void decoder_state_destroy(mw_model*) {}
}
