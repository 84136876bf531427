/* ORACLE (test infrastructure, not product code) — plain-C restatement of the log-mel front end.
 *
 * Follows whisperx.audio.log_mel_spectrogram(audio, n_mels, padding) as specified in SURVEY.md Appendix A.3 (whisperx
 * 3.7.6 is the reference's only pin, /root/reference/transcribe_colab.ipynb:47,80; reached through
 * /root/reference/transcribe.py:123) and librosa.filters.mel(sr=16000, n_fft=400, n_mels) (slaney scale + norm) for the
 * filterbank.  Deliberately naive and independent of torch: double precision, direct 400-point DFT per frame.
 * Pinned by tests/test_oracle_logmel.py against the torch oracle, the Hugging Face twin and tests/golden/.
 *
 * usage: logmel_ref <audio.f32> <n_samples> <padding> <n_mels> <out.f32>
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define N_FFT 400
#define HOP 160
#define N_FREQ 201
#define SR 16000.0

static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 1000.0 / (200.0 / 3.0), logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 1000.0 / (200.0 / 3.0), logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: %s audio.f32 n padding n_mels out.f32\n", argv[0]); return 2; }
    const long n = atol(argv[2]), padding = atol(argv[3]);
    const int n_mels = atoi(argv[4]);
    const long T = n + padding, n_frames = T / HOP;
    if (T <= N_FFT / 2) { fprintf(stderr, "input too short for reflect padding\n"); return 3; }
    float* a = (float*)calloc((size_t)(n > 0 ? n : 1), sizeof(float));
    FILE* f = fopen(argv[1], "rb");
    if (!f || (n > 0 && fread(a, sizeof(float), (size_t)n, f) != (size_t)n)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    /* filterbank [n_mels][201], float32 like the asset file */
    float* filt = (float*)calloc((size_t)n_mels * N_FREQ, sizeof(float));
    double* pts = (double*)malloc(sizeof(double) * (n_mels + 2));
    const double mmax = hz_to_mel(SR / 2.0);
    for (int i = 0; i < n_mels + 2; ++i) pts[i] = mel_to_hz(mmax * i / (n_mels + 1));
    for (int m = 0; m < n_mels; ++m) {
        const double enorm = 2.0 / (pts[m + 2] - pts[m]);
        for (int k = 0; k < N_FREQ; ++k) {
            const double fk = (SR / 2.0) * k / (N_FREQ - 1);
            const double lower = (fk - pts[m]) / (pts[m + 1] - pts[m]);
            const double upper = (pts[m + 2] - fk) / (pts[m + 2] - pts[m + 1]);
            double w = lower < upper ? lower : upper;
            if (w < 0) w = 0;
            filt[m * N_FREQ + k] = (float)(w * enorm);
        }
    }
    double win[N_FFT], cs[N_FFT], sn[N_FFT];
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < N_FFT; ++i) {
        win[i] = 0.5 - 0.5 * cos(2.0 * PI * i / N_FFT);   /* periodic Hann */
        cs[i] = cos(2.0 * PI * i / N_FFT);
        sn[i] = sin(2.0 * PI * i / N_FFT);
    }
    double* logspec = (double*)malloc(sizeof(double) * (size_t)n_mels * n_frames);
    double gmax = -1e300;
    double frame[N_FFT], power[N_FREQ];
    for (long t = 0; t < n_frames; ++t) {
        for (int i = 0; i < N_FFT; ++i) {
            long j = t * HOP + i - N_FFT / 2;                 /* center=True */
            if (j < 0) j = -j;                                /* reflect, edge sample not repeated */
            else if (j >= T) j = 2 * (T - 1) - j;
            frame[i] = (j >= 0 && j < n ? (double)a[j] : 0.0) * win[i];   /* zero padding beyond n */
        }
        for (int k = 0; k < N_FREQ; ++k) {
            double re = 0.0, im = 0.0;
            for (int i = 0; i < N_FFT; ++i) {
                const int idx = (int)(((long)k * i) % N_FFT);
                re += frame[i] * cs[idx];
                im -= frame[i] * sn[idx];
            }
            power[k] = re * re + im * im;
        }
        for (int m = 0; m < n_mels; ++m) {
            double acc = 0.0;
            for (int k = 0; k < N_FREQ; ++k) acc += (double)filt[m * N_FREQ + k] * power[k];
            const double v = log10(acc > 1e-10 ? acc : 1e-10);
            logspec[(size_t)m * n_frames + t] = v;
            if (v > gmax) gmax = v;
        }
    }
    float* out = (float*)malloc(sizeof(float) * (size_t)n_mels * n_frames);
    for (size_t i = 0; i < (size_t)n_mels * n_frames; ++i) {
        double v = logspec[i] > gmax - 8.0 ? logspec[i] : gmax - 8.0;
        out[i] = (float)((v + 4.0) / 4.0);
    }
    f = fopen(argv[5], "wb");
    fwrite(out, sizeof(float), (size_t)n_mels * n_frames, f);
    fclose(f);
    return 0;
}
