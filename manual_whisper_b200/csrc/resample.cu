// Audio decode on the GPU (replaces the `ffmpeg -ac 1 -ar 16000 -f s16le` pipe behind
// whisperx.load_audio, /root/reference/transcribe.py:117, for PCM input): interleaved int16 / float32 PCM at any rate and
// channel count -> mono float32 at 16 kHz.  Channel mean, then the polyphase windowed-sinc of
// torchaudio.functional.resample (taps built on the host, manual_whisper_b200/audio.py: sinc_resample_kernel), then the
// optional s16 quantisation the pipe implies.  One thread per output sample; a phase only walks its non-zero taps.
// HBM-bound by bytes (2*channels*orig/new bytes in, 4 out per output sample); the overlapping windows of neighbouring
// threads are served by L1/L2.
#include "mw_common.cuh"
#include <stdlib.h>

namespace mw {
namespace {

template <typename T> __device__ __forceinline__ float pcm_to_float(T v);
template <> __device__ __forceinline__ float pcm_to_float<int16_t>(int16_t v) { return (float)v * (1.0f / 32768.0f); }
template <> __device__ __forceinline__ float pcm_to_float<float>(float v) { return v; }

template <typename T>
__global__ void __launch_bounds__(256)
pcm_resample_kernel(const T* __restrict__ pcm, int64_t n_frames, int channels, int orig, int nw,
                    const float* __restrict__ kernels, const int* __restrict__ lo_hi, int taps, int width,
                    float* __restrict__ out, int64_t n_out, int quantize) {
    const int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (m >= n_out) return;
    const int64_t j = m / nw;
    const int i = (int)(m - j * nw);
    const int lo = lo_hi[2 * i], hi = lo_hi[2 * i + 1];
    const float* kr = kernels + (int64_t)i * taps;
    const int64_t start = j * orig - width;
    const float inv_c = 1.0f / (float)channels;
    float acc = 0.0f;
    for (int k = lo; k < hi; ++k) {
        const int64_t f = start + k;
        if (f < 0 || f >= n_frames) continue;
        float x;
        if (channels == 1) {
            x = pcm_to_float<T>(pcm[f]);
        } else {
            float s = 0.0f;
            for (int c = 0; c < channels; ++c) s += pcm_to_float<T>(pcm[f * channels + c]);
            x = s * inv_c;
        }
        acc = fmaf(__ldg(kr + k), x, acc);
    }
    if (quantize) acc = fminf(fmaxf(rintf(acc * 32768.0f), -32768.0f), 32767.0f) * (1.0f / 32768.0f);
    out[m] = acc;
}

// Tiled form: the 256 outputs of a CTA need one contiguous span of input frames; it is mixed to mono once into shared
// memory (coalesced loads, each frame read and converted once instead of ~13 times out of L1) and the taps run over it.
// Same arithmetic in the same order as the direct kernel, so results are identical.
template <typename T>
__global__ void __launch_bounds__(256)
pcm_resample_tiled_kernel(const T* __restrict__ pcm, int64_t n_frames, int channels, int orig, int nw,
                          const float* __restrict__ kernels, const int* __restrict__ lo_hi, int taps, int width,
                          float* __restrict__ out, int64_t n_out, int quantize) {
    extern __shared__ float tile[];
    const int64_t m0 = (int64_t)blockIdx.x * 256;
    const int64_t m1 = min(m0 + 256, n_out) - 1;
    const int64_t f_lo = (m0 / nw) * orig - width;
    const int span = (int)((m1 / nw) * orig - width + taps - f_lo);
    const float inv_c = 1.0f / (float)channels;
    for (int i = threadIdx.x; i < span; i += 256) {
        const int64_t f = f_lo + i;
        float x = 0.0f;
        if (f >= 0 && f < n_frames) {
            if (channels == 1) {
                x = pcm_to_float<T>(pcm[f]);
            } else {
                float sum = 0.0f;
                for (int c = 0; c < channels; ++c) sum += pcm_to_float<T>(pcm[f * channels + c]);
                x = sum * inv_c;
            }
        }
        tile[i] = x;
    }
    __syncthreads();
    const int64_t m = m0 + threadIdx.x;
    if (m >= n_out) return;
    const int64_t j = m / nw;
    const int i = (int)(m - j * nw);
    const int lo = lo_hi[2 * i], hi = lo_hi[2 * i + 1];
    const float* kr = kernels + (int64_t)i * taps;
    const float* x = tile + (int)(j * orig - width - f_lo);
    float acc = 0.0f;
    for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(kr + k), x[k], acc);
    if (quantize) acc = fminf(fmaxf(rintf(acc * 32768.0f), -32768.0f), 32767.0f) * (1.0f / 32768.0f);
    out[m] = acc;
}

}  // namespace
}  // namespace mw

extern "C" mw_status mw_pcm_resample(const void* d_pcm, int64_t n_frames, int channels, int sample_format, int orig, int new_rate,
                                     const float* d_kernels, const int32_t* d_lo_hi, int taps, int width, float* d_out,
                                     int64_t n_out, int quantize_s16, void* stream) {
    using namespace mw;
    MW_REQUIRE(d_pcm && d_kernels && d_lo_hi && d_out, "mw_pcm_resample: null argument");
    MW_REQUIRE(n_frames >= 0 && channels >= 1 && channels <= 8, "mw_pcm_resample: channels must be 1..8");
    MW_REQUIRE(sample_format == 0 || sample_format == 1, "mw_pcm_resample: sample_format must be 0 (s16) or 1 (f32)");
    MW_REQUIRE(orig >= 1 && new_rate >= 1 && taps == 2 * width + orig, "mw_pcm_resample: taps must equal 2*width + orig");
    MW_REQUIRE(n_out >= 0 && n_out <= (n_frames * new_rate + orig - 1) / orig, "mw_pcm_resample: n_out exceeds ceil(n_frames*new/orig)");
    if (n_out == 0) return MW_OK;
    const unsigned grid = (unsigned)((n_out + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t span_max = (int64_t)(255 / new_rate + 1) * orig + taps;        // input frames one CTA's outputs touch
    static const bool direct_only = [] { const char* e = getenv("MW_RESAMPLE_DIRECT"); return e && e[0] == '1'; }();   // A/B hook
    if (span_max * 4 <= 48 * 1024 && !direct_only) {
        const size_t smem = (size_t)span_max * 4;
        if (sample_format == 0)
            pcm_resample_tiled_kernel<int16_t><<<grid, 256, smem, st>>>((const int16_t*)d_pcm, n_frames, channels, orig, new_rate,
                                                                        d_kernels, d_lo_hi, taps, width, d_out, n_out, quantize_s16);
        else
            pcm_resample_tiled_kernel<float><<<grid, 256, smem, st>>>((const float*)d_pcm, n_frames, channels, orig, new_rate,
                                                                      d_kernels, d_lo_hi, taps, width, d_out, n_out, quantize_s16);
        MW_LAUNCH_CHECK();
        return MW_OK;
    }
    if (sample_format == 0)
        pcm_resample_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t*)d_pcm, n_frames, channels, orig, new_rate, d_kernels,
                                                            d_lo_hi, taps, width, d_out, n_out, quantize_s16);
    else
        pcm_resample_kernel<float><<<grid, 256, 0, st>>>((const float*)d_pcm, n_frames, channels, orig, new_rate, d_kernels,
                                                          d_lo_hi, taps, width, d_out, n_out, quantize_s16);
    MW_LAUNCH_CHECK();
    return MW_OK;
}
