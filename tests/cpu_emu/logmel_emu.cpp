// CPU emulation of the log-mel CTA (no GPU needed): runs the very stage functions of
// manual_whisper_b200/csrc/logmel_core.cuh with a loop over `tid` between barriers, so the index math,
// the FFT factorisation and the padding rules can be checked against the oracle in the CPU test-suite.
//
// usage: logmel_emu <audio.f32> <len> <padded> <n_mels> <filters.f32> <out.f32>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../manual_whisper_b200/csrc/logmel_core.cuh"

using namespace mw::logmel;

static std::vector<float> read_f32(const char* path, size_t n) {
    std::vector<float> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), 4, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc != 7) { fprintf(stderr, "usage\n"); return 2; }
    const long len = atol(argv[2]), padded = atol(argv[3]);
    const int n_mels = atoi(argv[4]);
    std::vector<float> audio = read_f32(argv[1], (size_t)(len > 0 ? len : 0));
    if (audio.empty()) audio.push_back(0.f);
    std::vector<float> filt = read_f32(argv[5], (size_t)n_mels * N_FREQ);
    const long n_frames = padded / HOP;
    // tables exactly as mw_logmel_plan_create builds them
    std::vector<float> win(N_FFT);
    std::vector<cpx> tw200(200), tw400(N_FREQ);
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < N_FFT; ++n) win[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / N_FFT));
    for (int m = 0; m < 200; ++m) tw200[m] = {(float)cos(2.0 * PI * m / 200.0), (float)(-sin(2.0 * PI * m / 200.0))};
    for (int k = 0; k <= 200; ++k) tw400[k] = {(float)cos(2.0 * PI * k / 400.0), (float)(-sin(2.0 * PI * k / 400.0))};
    std::vector<int> lo(n_mels), cnt(n_mels), off(n_mels);
    std::vector<float> w;
    for (int m = 0; m < n_mels; ++m) {
        int a = N_FREQ, b = -1;
        for (int k = 0; k < N_FREQ; ++k)
            if (filt[m * N_FREQ + k] != 0.0f) { if (k < a) a = k; b = k; }
        lo[m] = (b >= 0) ? a : 0;
        cnt[m] = (b >= 0) ? (b - a + 1) : 0;
        off[m] = (int)w.size();
        for (int k = 0; k < cnt[m]; ++k) w.push_back(filt[m * N_FREQ + lo[m] + k]);
    }
    if (w.empty()) w.push_back(0.f);
    std::vector<float> out((size_t)n_mels * n_frames, 0.f);
    std::vector<cpx> Y(FR * 200);
    std::vector<float> stage(STAGE_N), P(FR * PS);
    float gmax = -INFINITY;
    for (long f0 = 0; f0 < n_frames; f0 += FR) {
        if (tile_is_interior(audio.data(), len, padded, f0))
            for (int t = 0; t < NT; ++t) stage_load_fast(t, stage.data(), audio.data() + (f0 * HOP - N_FFT / 2));
        else
            for (int t = 0; t < NT; ++t) stage_load(t, stage.data(), audio.data(), len, padded, f0);
        for (int t = 0; t < NT; ++t) stage_radix8(t, stage.data(), win.data(), tw200.data(), Y.data());
        for (int t = 0; t < NT; ++t) stage_radix25(t, Y.data());
        for (int t = 0; t < NT; ++t) stage_power(t, Y.data(), tw400.data(), P.data());
        for (int t = 0; t < NT; ++t)
            gmax = fmaxf(gmax, stage_mel(t, P.data(), n_mels, lo.data(), cnt.data(), off.data(), w.data(), out.data(),
                                         n_frames, f0, n_frames, -INFINITY));
    }
    for (auto& v : out) v = finalize_value(v, gmax);
    FILE* f = fopen(argv[6], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    return 0;
}
