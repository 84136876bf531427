// Device-side VAD front end: frame energies -> speech turns -> <= 30 s windows, all on the GPU, so
// the waveform is uploaded once and never revisited by the host; only the finished window table (a few hundred numbers)
// goes back.  It stands where whisperx runs the pyannote segmentation network + Binarize(onset, offset, max_duration) +
// Vad.merge_chunks (knobs from /root/reference/transcribe.py:43-46,112).  The network's weights are not available
// offline, so the per-frame speech score is an ENERGY score (documented stand-in, manual_whisper_b200/vad.py: EnergyVad is
// the exact host twin, same arithmetic in float64); Binarize's hysteresis, the gap fill / blip removal and merge_chunks
// are the published algorithms.
//
//   1. vad_hist_kernel      : dB of every frame RMS -> 2048-bin histogram over [-140, 20] dB (integer counts: deterministic)
//   2. vad_segment_kernel   : one CTA: percentiles (10 %, 95 %) from the histogram -> score in [0, 1]; thread 0 walks the
//                             frames (scores staged through shared memory by the whole CTA) with the onset/offset hysteresis
//                             and the maximum turn duration, fills short gaps, drops blips, then runs merge_chunks and writes
//                             the window table
#include "mw_common.cuh"
#include <algorithm>

namespace mw {
namespace {

constexpr int VAD_BINS = 2048;
constexpr double VAD_DB_LO = -140.0, VAD_DB_HI = 20.0;
constexpr int VAD_TILE = 4096;

__device__ __forceinline__ double frame_db(float rms) { return 20.0 * log10((double)rms); }
__device__ __forceinline__ int db_bin(double db) {
    const int b = (int)floor((db - VAD_DB_LO) * (VAD_BINS / (VAD_DB_HI - VAD_DB_LO)));
    return b < 0 ? 0 : (b >= VAD_BINS ? VAD_BINS - 1 : b);
}

__global__ void vad_hist_kernel(const float* __restrict__ rms, int64_t n, unsigned int* __restrict__ hist) {
    __shared__ unsigned int h[VAD_BINS];
    for (int i = threadIdx.x; i < VAD_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&h[db_bin(frame_db(rms[i]))], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < VAD_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], h[i]);
}

struct VadParams {
    double onset, offset, max_duration, frame_s, min_on, min_off, chunk_size;
    int max_turns, max_windows;
};

// lower edge of the bin in which the cumulative count first exceeds q * n
__device__ double hist_percentile(const unsigned int* hist, int64_t n, double q) {
    const double target = q * (double)n;
    unsigned long long cum = 0;
    for (int b = 0; b < VAD_BINS; ++b) {
        cum += hist[b];
        if ((double)cum > target) return VAD_DB_LO + b * ((VAD_DB_HI - VAD_DB_LO) / VAD_BINS);
    }
    return VAD_DB_HI;
}

__global__ void __launch_bounds__(256)
vad_segment_kernel(const float* __restrict__ rms, int64_t n, const unsigned int* __restrict__ hist, VadParams p,
                   double* __restrict__ turns /* [max_turns][2] scratch */, double* __restrict__ windows /* [max_windows][2] */,
                   int* __restrict__ counts /* [0] turns, [1] windows, [2] overflow flag */) {
    __shared__ float tile[VAD_TILE];
    __shared__ double s_lo, s_span;
    if (threadIdx.x == 0) {
        const double lo = hist_percentile(hist, n, 0.10), hi = hist_percentile(hist, n, 0.95);
        s_lo = lo;
        s_span = fmax(hi - lo, 6.0);
    }
    __syncthreads();
    // ---- hysteresis over the frames (Binarize: onset / offset thresholds, turns cut at max_duration)
    int n_turns = 0, overflow = 0;
    bool active = false;
    double start = 0.0;
    for (int64_t base = 0; base < n; base += VAD_TILE) {
        const int m = (int)min((int64_t)VAD_TILE, n - base);
        for (int i = threadIdx.x; i < m; i += blockDim.x) tile[i] = rms[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < m; ++i) {
                double s = (frame_db(tile[i]) - s_lo) / s_span;
                s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
                const double t = (double)(base + i) * p.frame_s;
                if (!active && s > p.onset) { active = true; start = t; }
                else if (active && (s < p.offset || t - start >= p.max_duration)) {
                    if (n_turns < p.max_turns) { turns[2 * n_turns] = start; turns[2 * n_turns + 1] = t; ++n_turns; } else overflow = 1;
                    active = s >= p.offset;
                    start = t;
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    if (active) {
        if (n_turns < p.max_turns) { turns[2 * n_turns] = start; turns[2 * n_turns + 1] = (double)n * p.frame_s; ++n_turns; } else overflow = 1;
    }
    // ---- fill short gaps, drop blips (in place: the write index never passes the read index)
    int w = 0;
    for (int i = 0; i < n_turns; ++i) {
        const double a = turns[2 * i], b = turns[2 * i + 1];
        if (w > 0 && a - turns[2 * (w - 1) + 1] < p.min_off && b - turns[2 * (w - 1)] <= p.max_duration) turns[2 * (w - 1) + 1] = b;
        else { turns[2 * w] = a; turns[2 * w + 1] = b; ++w; }
    }
    int kept = 0;
    for (int i = 0; i < w; ++i)
        if (turns[2 * i + 1] - turns[2 * i] >= p.min_on) { turns[2 * kept] = turns[2 * i]; turns[2 * kept + 1] = turns[2 * i + 1]; ++kept; }
    counts[0] = kept;
    // ---- Vad.merge_chunks: greedy left-to-right merge into windows of at most chunk_size seconds
    int n_win = 0;
    if (kept > 0) {
        double curr_start = turns[0], curr_end = 0.0;
        for (int i = 0; i < kept; ++i) {
            const double s = turns[2 * i], e = turns[2 * i + 1];
            if (e - curr_start > p.chunk_size && curr_end - curr_start > 0.0) {
                if (n_win < p.max_windows) { windows[2 * n_win] = curr_start; windows[2 * n_win + 1] = curr_end; ++n_win; } else overflow = 1;
                curr_start = s;
            }
            curr_end = e;
        }
        if (n_win < p.max_windows) { windows[2 * n_win] = curr_start; windows[2 * n_win + 1] = curr_end; ++n_win; } else overflow = 1;
    }
    counts[1] = n_win;
    counts[2] = overflow;
}

}  // namespace
}  // namespace mw

// d_rms: frame RMS as written by mw_frame_rms (n_frames values).  d_scratch: >= 2048 * 4 + max_turns * 16 bytes.
// d_windows: double [max_windows][2] (start_s, end_s); d_counts: int32 [3] = {turns kept, windows, overflow flag}.
extern "C" mw_status mw_vad_windows(const float* d_rms, int64_t n_frames, double frame_s, double onset, double offset,
                                    double max_duration_s, double min_on_s, double min_off_s, double chunk_size_s,
                                    void* d_scratch, int max_turns, double* d_windows, int max_windows, int32_t* d_counts,
                                    void* stream) {
    MW_REQUIRE(d_rms && d_scratch && d_windows && d_counts, "mw_vad_windows: null argument");
    MW_REQUIRE(n_frames >= 0 && frame_s > 0 && max_turns > 0 && max_windows > 0 && chunk_size_s > 0, "mw_vad_windows: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* hist = reinterpret_cast<unsigned int*>(d_scratch);
    double* turns = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_scratch) + mw::VAD_BINS * 4);
    MW_CUDA_CHECK(cudaMemsetAsync(hist, 0, mw::VAD_BINS * 4, st));
    MW_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 3 * sizeof(int32_t), st));
    if (n_frames == 0) return MW_OK;
    const int grid = (int)std::min<int64_t>(148, (n_frames + 255) / 256);
    mw::vad_hist_kernel<<<grid, 256, 0, st>>>(d_rms, n_frames, hist);
    MW_LAUNCH_CHECK();
    mw::VadParams p{onset, offset, max_duration_s, frame_s, min_on_s, min_off_s, chunk_size_s, max_turns, max_windows};
    mw::vad_segment_kernel<<<1, 256, 0, st>>>(d_rms, n_frames, hist, p, turns, d_windows, d_counts);
    MW_LAUNCH_CHECK();
    return MW_OK;
}
