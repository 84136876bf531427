// Fused log-mel kernels for sm_100a — replaces whisperx.audio.log_mel_spectrogram
// (reached from /root/reference/transcribe.py:123).
//
// Kernel 1 (logmel_tile_kernel): one CTA = 32 frames of one chunk.  Samples are staged once in shared
// memory (coalesced reads, reflect/zero padding resolved on the fly), each frame is windowed and
// transformed by an in-shared-memory 200-point complex FFT (radix 8 x 25), split to the 201 real-input
// bins, squared, projected through the sparse mel filterbank, log10-clamped and written; the per-chunk
// maximum is reduced CTA-wide and folded into one atomicMax per CTA.
// Kernel 2 (logmel_finalize_kernel): max(x, chunkmax-8), (x+4)/4 in place (the tile kernel's output is
// still L2-resident) and, optionally, the bf16 time-major copy the encoder's conv stem reads.
#include "mw_common.cuh"
#include "logmel_core.cuh"

#include <algorithm>
#include <vector>

using namespace mw::logmel;

namespace {

constexpr int MAX_SM_MELS = 128;
constexpr int MAX_SM_W = 1536;

struct TileSmem {
    cpx Y[FR * 200];          // 51200 B
    float stage[STAGE_N];     // 21440 B
    float P[FR * PS];         // 25728 B
    float win[N_FFT];         // 1600 B
    cpx tw200[200];           // 1600 B
    cpx tw400[N_FREQ + 1];    // 1616 B
    float red[NT / 32];
    float red_min[NT / 32];
    // sparse filterbank, resident for the CTA's lifetime (a warp walks one mel row at a time: keeping these in
    // shared memory removes a chain of dependent global loads per row)
    int mel_lo[MAX_SM_MELS], mel_cnt[MAX_SM_MELS], mel_off[MAX_SM_MELS];
    float mel_w[MAX_SM_W];
};

__device__ __forceinline__ unsigned ordered_bits(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned b) {
    return __uint_as_float((b & 0x80000000u) ? (b & 0x7fffffffu) : ~b);
}

// Persistent: grid = min(#tiles, 2 x SMs); a CTA loads the constant tables once and then walks tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ... (tile t = 32 frames of chunk t / tiles_per_chunk).
__global__ void __launch_bounds__(NT, 2)
logmel_tile_kernel(const float* __restrict__ audio, int64_t n_audio,
                   const int64_t* __restrict__ offsets, const int32_t* __restrict__ lengths,
                   int64_t single_len, int64_t padded, int64_t n_frames, int n_mels, int n_chunks,
                   const float* __restrict__ tables,   // win[400] | tw200[200*2] | tw400[202*2]
                   const int* __restrict__ mel_lo, const int* __restrict__ mel_cnt,
                   const int* __restrict__ mel_off, const float* __restrict__ mel_w, int mel_nnz,
                   float* __restrict__ out, unsigned* __restrict__ gmax, float* __restrict__ tile_min /* null: store unscaled */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& s = *reinterpret_cast<TileSmem*>(smem_raw);
    const int tid = threadIdx.x;
    // constant tables -> shared (divergent indices would serialise in the constant cache)
    for (int i = tid; i < N_FFT; i += NT) s.win[i] = tables[i];
    {
        const float2* t2 = reinterpret_cast<const float2*>(tables + N_FFT);
        for (int i = tid; i < 200; i += NT) { float2 v = t2[i]; s.tw200[i] = {v.x, v.y}; }
        for (int i = tid; i < N_FREQ; i += NT) { float2 v = t2[200 + i]; s.tw400[i] = {v.x, v.y}; }
    }
    const bool sm_tables = n_mels <= MAX_SM_MELS && mel_nnz <= MAX_SM_W;
    if (sm_tables) {
        for (int i = tid; i < n_mels; i += NT) { s.mel_lo[i] = mel_lo[i]; s.mel_cnt[i] = mel_cnt[i]; s.mel_off[i] = mel_off[i]; }
        for (int i = tid; i < mel_nnz; i += NT) s.mel_w[i] = mel_w[i];
    }
    const int* p_lo = sm_tables ? s.mel_lo : mel_lo;
    const int* p_cnt = sm_tables ? s.mel_cnt : mel_cnt;
    const int* p_off = sm_tables ? s.mel_off : mel_off;
    const float* p_w = sm_tables ? s.mel_w : mel_w;
    const int64_t tiles_per_chunk = (n_frames + FR - 1) / FR;
    const int64_t n_tiles = tiles_per_chunk * n_chunks;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int chunk = (int)(t / tiles_per_chunk);
        const int64_t frame0 = (t - chunk * tiles_per_chunk) * FR;
        int64_t off = 0, len = single_len;
        if (offsets) {
            off = offsets[chunk];
            len = lengths[chunk];
            if (off < 0) off = 0;
            if (off > n_audio) off = n_audio;
            if (len > n_audio - off) len = n_audio - off;
            if (len > padded) len = padded;
            if (len < 0) len = 0;
        }
        if (tile_is_interior(audio + off, len, padded, frame0)) stage_load_fast(tid, s.stage, audio + off + (frame0 * HOP - N_FFT / 2));
        else stage_load(tid, s.stage, audio + off, len, padded, frame0);
        __syncthreads();
        stage_radix8(tid, s.stage, s.win, s.tw200, s.Y);
        __syncthreads();
        stage_radix25(tid, s.Y);
        __syncthreads();
        stage_power(tid, s.Y, s.tw400, s.P);
        __syncthreads();
        float vmin = INFINITY;
        float vmax = stage_mel(tid, s.P, n_mels, p_lo, p_cnt, p_off, p_w, out + (int64_t)chunk * n_mels * n_frames, n_frames,
                               frame0, n_frames, -INFINITY, tile_min ? &vmin : nullptr);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        }
        if ((tid & 31) == 0) { s.red[tid >> 5] = vmax; s.red_min[tid >> 5] = vmin; }
        __syncthreads();
        if (tid == 0) {
            float m = s.red[0], mn = s.red_min[0];
#pragma unroll
            for (int w = 1; w < NT / 32; ++w) { m = fmaxf(m, s.red[w]); mn = fminf(mn, s.red_min[w]); }
            atomicMax(gmax + chunk, ordered_bits(m));
            if (tile_min) tile_min[t] = mn;
        }
        // the next tile's first shared-memory writes (stage) are separated from this tile's last reads (P, red) by
        // the barrier above and the one after its load stage
    }
}

// grid: (ceil(n_frames/32), n_chunks); block 256 (8 warps, warp w handles mels w, w+8, ...)
__global__ void __launch_bounds__(256)
logmel_finalize_kernel(float* __restrict__ out, const unsigned* __restrict__ gmax, int n_mels, int64_t n_frames,
                       mw_h* __restrict__ out_t /* [chunks, n_frames+2, n_mels] or null */) {
    __shared__ float tile[128][33];
    const int chunk = blockIdx.y;
    const int64_t f0 = (int64_t)blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float g = from_ordered_bits(gmax[chunk]);
    float* o = out + (int64_t)chunk * n_mels * n_frames;
    const bool valid = f0 + lane < n_frames;
    for (int m0 = 0; m0 < n_mels; m0 += 128) {
        for (int m = m0 + warp; m < min(n_mels, m0 + 128); m += 8) {
            float v = 0.0f;
            if (valid) {
                v = finalize_value(o[(int64_t)m * n_frames + f0 + lane], g);
                o[(int64_t)m * n_frames + f0 + lane] = v;
            }
            tile[m - m0][lane] = v;
        }
        if (out_t) {
            __syncthreads();
            mw_h* ot = out_t + (int64_t)chunk * (n_frames + 2) * n_mels;
            const int mcount = min(128, n_mels - m0);
            // 32 frames x mcount mels, mel fastest
            for (int i = threadIdx.x; i < 32 * mcount; i += 256) {
                const int f = i / mcount, m = i - f * mcount;
                if (f0 + f < n_frames) ot[(f0 + f + 1) * n_mels + m0 + m] = f2h(tile[m][f]);
            }
            __syncthreads();
        }
    }
    if (out_t && blockIdx.x == 0) {
        mw_h* ot = out_t + (int64_t)chunk * (n_frames + 2) * n_mels;
        for (int i = threadIdx.x; i < n_mels; i += 256) {
            ot[i] = f2h(0.0f);
            ot[(n_frames + 1) * n_mels + i] = f2h(0.0f);
        }
    }
}

// Un-chunked path: the tile kernel stored (v + 4) / 4 already and left every tile's minimum behind; only tiles holding a value
// below max - 8 have anything to clamp (noise-like audio: a handful per hour, by extreme-value statistics; digital silence: all
// of them).  One CTA per tile of 32 frames; the others return at once.
__global__ void __launch_bounds__(256)
logmel_clamp_scaled_kernel(float* __restrict__ out, const unsigned* __restrict__ gmax, const float* __restrict__ tile_min, int n_mels,
                           int64_t n_frames) {
    const float g = from_ordered_bits(gmax[0]);
    if (tile_min[blockIdx.x] >= g - 8.0f) return;
    const int64_t f0 = (int64_t)blockIdx.x * FR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (f0 + lane >= n_frames) return;
    for (int m = warp; m < n_mels; m += 8) {
        float* o = out + (int64_t)m * n_frames + f0 + lane;
        *o = clamp_scaled(*o, g);
    }
}

}  // namespace

struct mw_logmel_plan {
    int device = 0;
    int n_mels = 0;
    int max_chunks = 0;
    float* d_tables = nullptr;
    int* d_lo = nullptr;
    int* d_cnt = nullptr;
    int* d_off = nullptr;
    float* d_w = nullptr;
    int nnz = 0;
    int sm_count = 148;
    unsigned* d_gmax = nullptr;
    float* d_tile_min = nullptr;         // un-chunked path only: per-tile minimum, grown on demand
    int64_t tile_min_cap = 0;
};

extern "C" mw_status mw_logmel_plan_create(int n_mels, const float* h_filters, int max_chunks, int device,
                                           mw_logmel_plan** out_plan) {
    MW_REQUIRE(out_plan && h_filters, "mw_logmel_plan_create: null argument");
    MW_REQUIRE(n_mels > 0 && n_mels <= 1024, "mw_logmel_plan_create: n_mels=%d out of range", n_mels);
    MW_REQUIRE(max_chunks > 0 && max_chunks <= 65535, "mw_logmel_plan_create: max_chunks=%d out of range (1..65535)", max_chunks);
    mw::DeviceGuard guard(device);
    auto* p = new mw_logmel_plan();
    p->device = device;
    p->n_mels = n_mels;
    p->max_chunks = max_chunks;
    // twiddle / window tables in double, rounded once
    std::vector<float> tab(N_FFT + 2 * 200 + 2 * (N_FREQ + 1), 0.0f);
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < N_FFT; ++n) tab[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / N_FFT));
    for (int m = 0; m < 200; ++m) {
        tab[N_FFT + 2 * m] = (float)cos(2.0 * PI * m / 200.0);
        tab[N_FFT + 2 * m + 1] = (float)(-sin(2.0 * PI * m / 200.0));
    }
    for (int k = 0; k <= 200; ++k) {
        tab[N_FFT + 400 + 2 * k] = (float)cos(2.0 * PI * k / 400.0);
        tab[N_FFT + 400 + 2 * k + 1] = (float)(-sin(2.0 * PI * k / 400.0));
    }
    // sparse filter rows: [first nonzero, last nonzero]
    std::vector<int> lo(n_mels), cnt(n_mels), off(n_mels);
    std::vector<float> w;
    for (int m = 0; m < n_mels; ++m) {
        int a = N_FREQ, b = -1;
        for (int k = 0; k < N_FREQ; ++k)
            if (h_filters[m * N_FREQ + k] != 0.0f) { if (k < a) a = k; b = k; }
        lo[m] = (b >= 0) ? a : 0;
        cnt[m] = (b >= 0) ? (b - a + 1) : 0;
        off[m] = (int)w.size();
        for (int k = 0; k < cnt[m]; ++k) w.push_back(h_filters[m * N_FREQ + lo[m] + k]);
    }
    p->nnz = (int)w.size();
    if (w.empty()) w.push_back(0.0f);
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (p->sm_count <= 0) p->sm_count = 148;
    MW_CUDA_CHECK(cudaMalloc(&p->d_tables, tab.size() * sizeof(float)));
    MW_CUDA_CHECK(cudaMalloc(&p->d_lo, n_mels * sizeof(int)));
    MW_CUDA_CHECK(cudaMalloc(&p->d_cnt, n_mels * sizeof(int)));
    MW_CUDA_CHECK(cudaMalloc(&p->d_off, n_mels * sizeof(int)));
    MW_CUDA_CHECK(cudaMalloc(&p->d_w, w.size() * sizeof(float)));
    MW_CUDA_CHECK(cudaMalloc(&p->d_gmax, max_chunks * sizeof(unsigned)));
    MW_CUDA_CHECK(cudaMemcpy(p->d_tables, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    MW_CUDA_CHECK(cudaMemcpy(p->d_lo, lo.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
    MW_CUDA_CHECK(cudaMemcpy(p->d_cnt, cnt.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
    MW_CUDA_CHECK(cudaMemcpy(p->d_off, off.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
    MW_CUDA_CHECK(cudaMemcpy(p->d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    MW_CUDA_CHECK(cudaFuncSetAttribute(logmel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(TileSmem)));
    *out_plan = p;
    return MW_OK;
}

extern "C" void mw_logmel_plan_destroy(mw_logmel_plan* p) {
    if (!p) return;
    mw::DeviceGuard guard(p->device);
    cudaFree(p->d_tables); cudaFree(p->d_lo); cudaFree(p->d_cnt); cudaFree(p->d_off);
    cudaFree(p->d_w); cudaFree(p->d_gmax); cudaFree(p->d_tile_min);
    delete p;
}

static mw_status run_logmel(mw_logmel_plan* p, const float* d_audio, int64_t n_audio, const int64_t* d_offsets,
                            const int32_t* d_lengths, int n_chunks, int64_t single_len, int64_t padded,
                            float* d_out, void* d_out_t, cudaStream_t st) {
    const int64_t n_frames = padded / HOP;
    if (n_frames == 0 || n_chunks == 0) return MW_OK;
    const int64_t tiles = mw::ceil_div64(n_frames, FR);
    MW_REQUIRE(tiles <= 2147483647LL, "mw_logmel: clip too long");
    mw::DeviceGuard guard(p->device);
    MW_CUDA_CHECK(cudaMemsetAsync(p->d_gmax, 0, n_chunks * sizeof(unsigned), st));
    // one un-chunked clip with no time-major copy wanted (mw_logmel_long): values are stored scaled and the second pass over
    // the output runs only if the clamp has anything to do
    const bool scaled = !d_offsets && !d_out_t && n_chunks == 1;
    if (scaled && tiles > p->tile_min_cap) {          // first call at this length (the only allocation this path ever makes)
        MW_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(p->d_tile_min);
        p->d_tile_min = nullptr;
        p->tile_min_cap = 0;
        MW_CUDA_CHECK(cudaMalloc(&p->d_tile_min, tiles * sizeof(float)));
        p->tile_min_cap = tiles;
    }
    dim3 grid((unsigned)tiles, (unsigned)n_chunks);
    const int64_t total_tiles = tiles * n_chunks;
    const unsigned persistent = (unsigned)(total_tiles < 2LL * p->sm_count ? total_tiles : 2LL * p->sm_count);
    logmel_tile_kernel<<<persistent, NT, sizeof(TileSmem), st>>>(d_audio, n_audio, d_offsets, d_lengths, single_len, padded,
                                                                n_frames, p->n_mels, n_chunks, p->d_tables, p->d_lo, p->d_cnt,
                                                                p->d_off, p->d_w, p->nnz, d_out, p->d_gmax, scaled ? p->d_tile_min : nullptr);
    MW_LAUNCH_CHECK();
    if (scaled) {
        logmel_clamp_scaled_kernel<<<(unsigned)tiles, 256, 0, st>>>(d_out, p->d_gmax, p->d_tile_min, p->n_mels, n_frames);
    } else {
        logmel_finalize_kernel<<<grid, 256, 0, st>>>(d_out, p->d_gmax, p->n_mels, n_frames, (mw_h*)d_out_t);
    }
    MW_LAUNCH_CHECK();
    return MW_OK;
}

extern "C" mw_status mw_logmel(mw_logmel_plan* plan, const float* d_audio, int64_t n_audio, const int64_t* d_offsets,
                               const int32_t* d_lengths, int n_chunks, float* d_out, void* d_out_t, void* stream) {
    MW_REQUIRE(plan, "mw_logmel: null plan");
    MW_REQUIRE(n_chunks >= 0 && n_chunks <= plan->max_chunks, "mw_logmel: n_chunks=%d exceeds plan max_chunks=%d",
               n_chunks, plan->max_chunks);
    if (n_chunks == 0) return MW_OK;
    MW_REQUIRE(d_audio && d_offsets && d_lengths && d_out, "mw_logmel: null device pointer");
    return run_logmel(plan, d_audio, n_audio, d_offsets, d_lengths, n_chunks, 0, 480000, d_out, d_out_t,
                      (cudaStream_t)stream);
}

extern "C" mw_status mw_logmel_long(mw_logmel_plan* plan, const float* d_audio, int64_t n, int64_t padding,
                                    float* d_out, void* stream) {
    MW_REQUIRE(plan, "mw_logmel_long: null plan");
    MW_REQUIRE(n >= 0 && padding >= 0, "mw_logmel_long: negative length");
    // torch.stft(center=True, pad_mode='reflect') needs more than n_fft/2 samples
    MW_REQUIRE(n + padding > N_FFT / 2, "mw_logmel_long: input of %lld samples is too short for reflect padding of %d",
               (long long)(n + padding), N_FFT / 2);
    MW_REQUIRE(d_audio || n == 0, "mw_logmel_long: null audio");
    MW_REQUIRE(d_out, "mw_logmel_long: null output");
    return run_logmel(plan, d_audio, n, nullptr, nullptr, 1, n, n + padding, d_out, nullptr, (cudaStream_t)stream);
}
