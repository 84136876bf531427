// Log-mel front end, math core (host/device).  Replaces whisperx.audio.log_mel_spectrogram
// (reached from /root/reference/transcribe.py:123).
//
// Every stage is written as "what thread `tid` of an NT-thread CTA does between two barriers", so the
// same code runs inside the CUDA kernel (logmel.cu) and, compiled as plain C++, inside the CPU
// emulation harness tests/cpu_emu/logmel_emu.cpp that checks the index math without a GPU.
//
// A 400-point real frame is transformed as a 200-point complex FFT (z[n] = x[2n] + i x[2n+1]),
// 200 = 8 x 25: radix-8 butterflies (+ W200 twiddles) then in-register radix-25 (5 x 5), followed by
// the real-input split X[k] = E[k] + W400^k O[k], k = 0..200.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MW_HD __host__ __device__ __forceinline__
#else
#define MW_HD inline
#endif

namespace mw {
namespace logmel {

constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int N_FREQ = 201;
constexpr int FR = 32;                              // frames per CTA tile
#ifndef MW_LOGMEL_NT
#define MW_LOGMEL_NT 512
#endif
constexpr int NT = MW_LOGMEL_NT;                    // threads per CTA: 512 (64 registers, 2 CTAs/SM = 32 warps per SM) measured 5-10 % over 256; the radix-25 stage uses 256 of them
constexpr int STAGE_N = (FR - 1) * HOP + N_FFT;     // 5360 samples staged per tile
constexpr int PS = 201;                             // power row stride (odd: conflict-free across lanes)

struct alignas(8) cpx { float re, im; };
struct alignas(8) f2 { float x, y; };

MW_HD cpx cadd(cpx a, cpx b) { return {a.re + b.re, a.im + b.im}; }
MW_HD cpx csub(cpx a, cpx b) { return {a.re - b.re, a.im - b.im}; }
MW_HD cpx cmul(cpx a, cpx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
MW_HD cpx cscale(cpx a, float s) { return {a.re * s, a.im * s}; }
MW_HD cpx mul_mi(cpx a) { return {a.im, -a.re}; }   // a * (-i)
MW_HD cpx mul_pi(cpx a) { return {-a.im, a.re}; }   // a * (+i)

// forward DFT-5 in place (kernel exp(-2 pi i nk/5))
MW_HD void dft5(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4) {
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    cpx t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
    cpx x0 = cadd(a0, cadd(t1, t2));
    cpx m1 = {a0.re + c1 * t1.re + c2 * t2.re, a0.im + c1 * t1.im + c2 * t2.im};
    cpx m2 = {a0.re + c2 * t1.re + c1 * t2.re, a0.im + c2 * t1.im + c1 * t2.im};
    cpx n1 = {s1 * t3.re + s2 * t4.re, s1 * t3.im + s2 * t4.im};
    cpx n2 = {s2 * t3.re - s1 * t4.re, s2 * t3.im - s1 * t4.im};
    a0 = x0;
    a1 = cadd(m1, mul_mi(n1));
    a4 = cadd(m1, mul_pi(n1));
    a2 = cadd(m2, mul_mi(n2));
    a3 = cadd(m2, mul_pi(n2));
}

// forward DFT-8 in place, natural order in and out
MW_HD void dft8(cpx* a) {
    const float r = 0.70710678118654752f;
    // stage 1: pairs (j, j+4)
    cpx b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
    cpx b1 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
    cpx b2 = cadd(a[2], a[6]), b6 = csub(a[2], a[6]);
    cpx b3 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
    // even outputs: DFT-4 of (b0,b1,b2,b3)
    cpx c0 = cadd(b0, b2), c2 = csub(b0, b2);
    cpx c1 = cadd(b1, b3), c3 = mul_mi(csub(b1, b3));
    a[0] = cadd(c0, c1);
    a[4] = csub(c0, c1);
    a[2] = cadd(c2, c3);
    a[6] = csub(c2, c3);
    // odd outputs: DFT-4 of (b4, b5 W8^1, b6 W8^2, b7 W8^3)
    cpx d5 = {(b5.re + b5.im) * r, (b5.im - b5.re) * r};      // b5 * (1-i)/sqrt2
    cpx d6 = mul_mi(b6);                                      // b6 * (-i)
    cpx d7 = {(b7.im - b7.re) * r, -(b7.re + b7.im) * r};     // b7 * (-1-i)/sqrt2
    cpx e0 = cadd(b4, d6), e2 = csub(b4, d6);
    cpx e1 = cadd(d5, d7), e3 = mul_mi(csub(d5, d7));
    a[1] = cadd(e0, e1);
    a[5] = csub(e0, e1);
    a[3] = cadd(e2, e3);
    a[7] = csub(e2, e3);
}

// W25^m, m = 0..16 (products n2*k1 with n2,k1 in 0..4)
MW_HD cpx w25(int m) {
    const float c[17] = {1.0f, 0.96858316112863108f, 0.87630668004386358f, 0.72896862742141155f,
                         0.53582679497899666f, 0.30901699437494742f, 0.062790519529313374f,
                         -0.18738131458572463f, -0.42577929156507272f, -0.63742398974868975f,
                         -0.80901699437494742f, -0.92977648588825146f, -0.99211470131447788f,
                         -0.99211470131447788f, -0.92977648588825146f, -0.80901699437494742f,
                         -0.63742398974868975f};
    const float s[17] = {0.0f, 0.24868988716485479f, 0.48175367410171532f, 0.68454710592868873f,
                         0.84432792550201508f, 0.95105651629515357f, 0.99802672842827156f,
                         0.98228725072868872f, 0.90482705246601958f, 0.77051324277578925f,
                         0.58778525229247313f, 0.36812455268467797f, 0.12533323356430426f,
                         -0.12533323356430426f, -0.36812455268467797f, -0.58778525229247313f,
                         -0.77051324277578925f};
    return {c[m], -s[m]};
}

// forward DFT-25 in place: a[n], n = 5 n1 + n2  ->  a[k], k = k1 + 5 k2
MW_HD void dft25(cpx* a) {
    cpx b[25];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        cpx t0 = a[n2], t1 = a[5 + n2], t2 = a[10 + n2], t3 = a[15 + n2], t4 = a[20 + n2];
        dft5(t0, t1, t2, t3, t4);   // over n1 -> k1
        b[n2 * 5 + 0] = t0;
        b[n2 * 5 + 1] = n2 ? cmul(t1, w25(n2 * 1)) : t1;
        b[n2 * 5 + 2] = n2 ? cmul(t2, w25(n2 * 2)) : t2;
        b[n2 * 5 + 3] = n2 ? cmul(t3, w25(n2 * 3)) : t3;
        b[n2 * 5 + 4] = n2 ? cmul(t4, w25(n2 * 4)) : t4;
    }
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        cpx t0 = b[k1], t1 = b[5 + k1], t2 = b[10 + k1], t3 = b[15 + k1], t4 = b[20 + k1];
        dft5(t0, t1, t2, t3, t4);   // over n2 -> k2
        a[k1] = t0; a[k1 + 5] = t1; a[k1 + 10] = t2; a[k1 + 15] = t3; a[k1 + 20] = t4;
    }
}

// Sample j of the reflect-padded, zero-extended chunk.  `len` valid samples, `padded` = len + padding.
// torch.stft(center=True, pad_mode="reflect"): x[-k] = x[k], x[padded-1+k] = x[padded-1-k].
MW_HD float fetch_sample(const float* audio, int64_t j, int64_t len, int64_t padded) {
    if (j < 0) j = -j;
    else if (j >= padded) j = 2 * (padded - 1) - j;
    if (j < 0 || j >= len) return 0.0f;
    return audio[j];
}

// ---- stage 0: stage the tile's samples ------------------------------------------------------------
MW_HD void stage_load(int tid, float* stage, const float* audio, int64_t len, int64_t padded, int64_t frame0) {
    const int64_t j0 = frame0 * HOP - N_FFT / 2;
    for (int i = tid; i < STAGE_N; i += NT) stage[i] = fetch_sample(audio, j0 + i, len, padded);
}

// Interior tiles (no reflection, every sample inside the chunk): 16-byte global loads from the enclosing aligned
// range, scattered into the staging buffer with the sub-vector shift `mis` = (address of sample j0) mod 4 floats.
struct alignas(16) f4 { float x, y, z, w; };
MW_HD void stage_load_fast(int tid, float* stage, const float* first /* = audio + j0 */) {
    const int mis = (int)((reinterpret_cast<uintptr_t>(first) >> 2) & 3);
    const f4* base = reinterpret_cast<const f4*>(first - mis);
    const int nvec = (STAGE_N + mis + 3) >> 2;
    for (int v = tid; v < nvec; v += NT) {
        const f4 q = base[v];
        const int i = 4 * v - mis;
        if (i >= 0 && i < STAGE_N) stage[i] = q.x;
        if (i + 1 >= 0 && i + 1 < STAGE_N) stage[i + 1] = q.y;
        if (i + 2 >= 0 && i + 2 < STAGE_N) stage[i + 2] = q.z;
        if (i + 3 < STAGE_N) stage[i + 3] = q.w;
    }
}
// true when [j0, j0 + STAGE_N) rounded out to 16-byte vectors lies inside the chunk's valid samples
MW_HD bool tile_is_interior(const float* audio, int64_t len, int64_t padded, int64_t frame0) {
    const int64_t j0 = frame0 * HOP - N_FFT / 2;
    return j0 >= 4 && j0 + STAGE_N + 4 <= len && j0 + STAGE_N <= padded;
}

// ---- stage 1: window, radix-8 over n1, W200 twiddle ------------------------------------------------
// Y[f*200 + k1*25 + n2] = W200^(n2 k1) * sum_n1 z[25 n1 + n2] W8^(n1 k1)
MW_HD void stage_radix8(int tid, const float* stage, const float* win, const cpx* tw200, cpx* Y) {
    for (int q = tid; q < FR * 25; q += NT) {
        const int f = q / 25, n2 = q - f * 25;
        const float* x = stage + f * HOP;
        cpx a[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int idx = 2 * (25 * n1 + n2);
            const f2 xv = *reinterpret_cast<const f2*>(x + idx);     // idx and f*HOP are even: 8-byte aligned
            const f2 wv = *reinterpret_cast<const f2*>(win + idx);
            a[n1].re = xv.x * wv.x;
            a[n1].im = xv.y * wv.y;
        }
        dft8(a);
        cpx* y = Y + f * 200 + n2;
        y[0] = a[0];
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) y[k1 * 25] = cmul(a[k1], tw200[n2 * k1]);
    }
}

// ---- stage 2: radix-25 over n2, in place: slot [k1][k2] <- Z[k1 + 8 k2] ---------------------------
MW_HD void stage_radix25(int tid, cpx* Y) {
    // FR*8 tasks
    if (tid >= FR * 8) return;
    cpx* y = Y + (tid >> 3) * 200 + (tid & 7) * 25;
    cpx a[25];
#pragma unroll
    for (int j = 0; j < 25; ++j) a[j] = y[j];
    dft25(a);
#pragma unroll
    for (int j = 0; j < 25; ++j) y[j] = a[j];
}

MW_HD cpx z_at(const cpx* Yf, int k) { return Yf[(k & 7) * 25 + (k >> 3)]; }

// ---- stage 3: real-input split and power ----------------------------------------------------------
// One task handles the bin pair (k, 200-k): with E = (Z[k] + conj Z[200-k])/2, O = -i (Z[k] - conj Z[200-k])/2 and
// T = W400^k O, X[k] = E + T and X[200-k] = conj(E - T), so both powers come from one pair of loads.
MW_HD void stage_power(int tid, const cpx* Y, const cpx* tw400, float* P) {
    for (int q = tid; q < FR * 101; q += NT) {
        const int f = q / 101, k = q - f * 101;
        const cpx* Yf = Y + f * 200;
        const cpx zk = z_at(Yf, k);
        cpx zr = z_at(Yf, k == 0 ? 0 : 200 - k);
        zr.im = -zr.im;
        const cpx e = cscale(cadd(zk, zr), 0.5f);
        const cpx o = cscale(mul_mi(csub(zk, zr)), 0.5f);
        const cpx t = cmul(tw400[k], o);
        const cpx a = cadd(e, t), b = csub(e, t);
        P[f * PS + k] = a.re * a.re + a.im * a.im;
        P[f * PS + 200 - k] = b.re * b.re + b.im * b.im;     // k = 100 writes the same value twice
    }
}

// log10 through the hardware log2 (MUFU.LG2, ~2^-22 relative): abs error <= 3e-6 over the 1e-10..1e6 range used here
MW_HD float fast_log10(float x) {
#if defined(__CUDA_ARCH__)
    return __log2f(x) * 0.30102999566398120f;
#else
    return log2f(x) * 0.30102999566398120f;
#endif
}

// ---- stage 4: sparse mel projection + log10, one lane per frame -----------------------------------
// returns the thread's running max of the values it produced (for valid frames).  `vmin` (optional): running min as well, and
// the value is then stored already scaled, (v + 4) / 4 - the un-chunked path, whose clamp pass is skipped when min >= max - 8.
MW_HD float stage_mel(int tid, const float* P, int n_mels, const int* mel_lo, const int* mel_cnt,
                      const int* mel_off, const float* mel_w, float* out, int64_t out_stride,
                      int64_t frame0, int64_t n_frames, float vmax, float* vmin = nullptr) {
    const int lane = tid & 31, warp = tid >> 5;
    const float* p = P + lane * PS;
    const bool valid = frame0 + lane < n_frames;
    for (int m = warp; m < n_mels; m += NT / 32) {
        const int lo = mel_lo[m], cnt = mel_cnt[m];
        const float* w = mel_w + mel_off[m];
        float acc = 0.0f;
        for (int i = 0; i < cnt; ++i) acc = fmaf(w[i], p[lo + i], acc);
        const float v = fast_log10(fmaxf(acc, 1e-10f));
        if (valid) {
            out[(int64_t)m * out_stride + frame0 + lane] = vmin ? (v + 4.0f) / 4.0f : v;
            vmax = fmaxf(vmax, v);
            if (vmin) *vmin = fminf(*vmin, v);
        }
    }
    return vmax;
}

MW_HD float finalize_value(float v, float gmax) { return (fmaxf(v, gmax - 8.0f) + 4.0f) / 4.0f; }
// the same value from an already scaled w = (v + 4) / 4: x -> (x + 4) / 4 is monotone, so max commutes with it exactly
MW_HD float clamp_scaled(float w, float gmax) { return fmaxf(w, ((gmax - 8.0f) + 4.0f) / 4.0f); }

}  // namespace logmel
}  // namespace mw
