/*
 * oracle/features_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, f32) of the reference's log-mel frontend, /root/reference/rust/features/src/lib.rs.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 *
 * The Rust original cannot be built here (no cargo/rustc; realfft 3.5.0 / rustfft 6.4.1 from
 * rust/Cargo.lock:503,556 are not vendored).  Any exact f32 real FFT of size 512 computes the same
 * spectrum up to f32 rounding, so the FFT below is a textbook radix-2 with f64-derived twiddles.
 *
 * Parity pin: the reference holds ONE test for this path (lib.rs:229-241: 16000 zeros -> 98*128 values,
 * shape only).  tests/test_oracle_features.py checks that plus analytic known answers
 * (zeros -> ln(1e-5), pure tone -> peak mel bin, filterbank shape).  Value-level numerics vs the Rust
 * binary are otherwise UNPINNED (stated in DESIGN.md).
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define SR 16000
#define N_FFT 512
#define WIN 400
#define HOP 160
#define N_MELS 128
#define N_BINS (N_FFT / 2 + 1)

static float g_window[WIN];
static float g_fb[N_MELS][N_BINS];
static float g_tw_re[N_FFT / 2], g_tw_im[N_FFT / 2];
static int g_init = 0;

/* lib.rs:180-186 */
static float hz_to_mel(float hz) { return 2595.0f * log10f(1.0f + hz / 700.0f); }
static float mel_to_hz(float mel) { return 700.0f * (powf(10.0f, mel / 2595.0f) - 1.0f); }

static void init_tables(void) {
  if (g_init) return;
  /* lib.rs:174-178  symmetric Hann, f32 arithmetic: 0.5*(1-cos(2*pi*i/(size-1))) */
  for (int i = 0; i < WIN; ++i)
    g_window[i] = 0.5f * (1.0f - cosf(2.0f * 3.14159265358979323846f * (float)i / (float)(WIN - 1)));
  /* lib.rs:188-223  HTK-mel triangles, no area normalisation, f_min 0, f_max sr/2 */
  float min_mel = hz_to_mel(0.0f), max_mel = hz_to_mel((float)(SR / 2));
  float pts[N_MELS + 2];
  for (int i = 0; i < N_MELS + 2; ++i) {
    float mel = min_mel + (max_mel - min_mel) * ((float)i / (float)(N_MELS + 1));
    pts[i] = mel_to_hz(mel);
  }
  memset(g_fb, 0, sizeof(g_fb));
  for (int m = 0; m < N_MELS; ++m) {
    float left = pts[m], center = pts[m + 1], right = pts[m + 2];
    for (int i = 0; i < N_BINS; ++i) {
      float freq = (float)i * (float)SR / (float)N_FFT;
      if (freq > left && freq < center)
        g_fb[m][i] = (freq - left) / (center - left);
      else if (freq >= center && freq < right)
        g_fb[m][i] = (right - freq) / (right - center);
    }
  }
  for (int k = 0; k < N_FFT / 2; ++k) {
    double a = -2.0 * 3.14159265358979323846 * (double)k / (double)N_FFT;
    g_tw_re[k] = (float)cos(a);
    g_tw_im[k] = (float)sin(a);
  }
  g_init = 1;
}

/* in-place radix-2 DIT complex FFT, size 512, f32 */
static void fft512(float* re, float* im) {
  for (int i = 1, j = 0; i < N_FFT; ++i) {
    int bit = N_FFT >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      float t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  for (int len = 2; len <= N_FFT; len <<= 1) {
    int half = len >> 1, step = N_FFT / len;
    for (int i = 0; i < N_FFT; i += len)
      for (int k = 0; k < half; ++k) {
        float wr = g_tw_re[k * step], wi = g_tw_im[k * step];
        float xr = re[i + k + half], xi = im[i + k + half];
        float tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
        re[i + k + half] = re[i + k] - tr;
        im[i + k + half] = im[i + k] - ti;
        re[i + k] += tr;
        im[i + k] += ti;
      }
  }
}

/* lib.rs:66-120: number of frames = floor((n-400)/160)+1 for n>=400, else 0 */
size_t fr_num_frames(size_t n) { return n < WIN ? 0 : (n - WIN) / HOP + 1; }

static void one_frame(const float* frame, float* out128) {
  float re[N_FFT], im[N_FFT], pw[N_BINS];
  for (int i = 0; i < WIN; ++i) re[i] = frame[i] * g_window[i]; /* preemphasis = 0.0 (lib.rs:22) */
  for (int i = WIN; i < N_FFT; ++i) re[i] = 0.0f;               /* tail zero pad (lib.rs:88-90) */
  memset(im, 0, sizeof(im));
  fft512(re, im);
  for (int i = 0; i < N_BINS; ++i) pw[i] = re[i] * re[i] + im[i] * im[i];
  for (int m = 0; m < N_MELS; ++m) {
    float e = 0.0f;
    for (int i = 0; i < N_BINS; ++i) e += pw[i] * g_fb[m][i]; /* dense dot, bin order (lib.rs:103-107) */
    out128[m] = logf(e + 1e-5f);
  }
}

typedef struct { const float* audio; float* out; long t0, t1; } fr_job;
static void* fr_worker(void* p) {
  fr_job* j = (fr_job*)p;
  for (long t = j->t0; t < j->t1; ++t) one_frame(j->audio + (size_t)t * HOP, j->out + (size_t)t * N_MELS);
  return NULL;
}

/* audio[n] -> out[T*128] frames-major.  threads<=1: serial like the Rust original; else frames are split
 * over `threads` pthreads (frames are independent, so the values are identical to the serial run). */
void fr_logmel(const float* audio, size_t n, float* out, int threads) {
  init_tables();
  long T = (long)fr_num_frames(n);
  if (threads <= 1 || T < 4 * threads) {
    fr_job j = {audio, out, 0, T};
    fr_worker(&j);
    return;
  }
  if (threads > 256) threads = 256;
  pthread_t th[256];
  fr_job jobs[256];
  for (int i = 0; i < threads; ++i) {
    jobs[i].audio = audio; jobs[i].out = out;
    jobs[i].t0 = T * i / threads; jobs[i].t1 = T * (i + 1) / threads;
    pthread_create(&th[i], NULL, fr_worker, &jobs[i]);
  }
  for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
}

/* lib.rs:127-159 */
void fr_per_feature_stats(const float* f, size_t frames, float* mean, float* std) {
  for (int m = 0; m < N_MELS; ++m) mean[m] = std[m] = 0.0f;
  if (frames == 0) return;
  for (size_t t = 0; t < frames; ++t)
    for (int m = 0; m < N_MELS; ++m) mean[m] += f[t * N_MELS + m];
  float denom = (float)frames;
  for (int m = 0; m < N_MELS; ++m) mean[m] /= denom;
  float denom_std = frames > 1 ? (float)(frames - 1) : 1.0f;
  for (size_t t = 0; t < frames; ++t)
    for (int m = 0; m < N_MELS; ++m) {
      float d = f[t * N_MELS + m] - mean[m];
      std[m] += d * d;
    }
  for (int m = 0; m < N_MELS; ++m) std[m] = sqrtf(std[m] / denom_std) + 1e-5f;
}

/* lib.rs:161-172 */
void fr_apply_norm(float* f, size_t frames, const float* mean, const float* std) {
  for (size_t t = 0; t < frames; ++t)
    for (int m = 0; m < N_MELS; ++m) f[t * N_MELS + m] = (f[t * N_MELS + m] - mean[m]) / std[m];
}

/* rust/cli/src/main.rs:78-88: [T,128] -> [128,T] (the C-ABI layout) */
void fr_frames_to_bins_major(const float* tc, size_t frames, float* ct) {
  for (size_t t = 0; t < frames; ++t)
    for (int m = 0; m < N_MELS; ++m) ct[(size_t)m * frames + t] = tc[t * N_MELS + m];
}

void fr_get_tables(float* window400, float* fb_128x257) {
  init_tables();
  memcpy(window400, g_window, sizeof(g_window));
  memcpy(fb_128x257, g_fb, sizeof(g_fb));
}
