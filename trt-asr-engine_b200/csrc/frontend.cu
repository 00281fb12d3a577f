// frontend.cu -- log-mel frontend on sm_100a (piece 1 of the hot path).
//
// Replaces the CPU extractor /root/reference/rust/features/src/lib.rs:66-120 (framing, symmetric Hann-400,
// tail zero-pad to 512, real FFT, |X|^2, 128 un-normalised HTK-mel triangles, ln(E+1e-5)) and the
// per-feature normalisation :127-172.  HBM-bound by design: one pass reads 16000 samples and writes
// 100x128 features per audio second (115 200 B); everything in between lives in shared memory / registers.
//
// Mapping: one warp per frame.  The 512-point real FFT is done as a 256-point complex FFT
// (z[n] = x[2n] + i x[2n+1]) followed by the even/odd split.  The mel stage is sparse: each triangle
// touches a contiguous bin range, summed in bin order.
//   logmel_reg_kernel (default): the FFT lives in REGISTERS -- 256 = 8 x 8 x 4, every lane holds 8 points, three in-register
//     butterfly passes with two conflict-free transposes through the warp's shared-memory slice in between; window samples and
//     all twiddles are per-lane loop invariants; the even/odd split pairs Z[k] with Z[256-k] by one shuffle per point; mel
//     weights are stored transposed ([group][q][lane], zero-padded to the longest triangle of each group of 32 filters).
//     ~190 shared-memory wavefronts per frame instead of ~400: the round-1 kernel was bound by exactly that pipe.
//   logmel_kernel (PARAKEET_B200_LOGMEL=0): radix-4 Stockham autosort in shared memory (4 stages, 2 radix-4 butterflies per lane
//     per stage); kept as the second implementation the register kernel is tested against.
#include "frontend.h"

#include <math.h>
#include <stdlib.h>

#include <vector>

namespace pkb {

constexpr int kWarpsPerCta = 8;
constexpr int kNfft = 512, kWin = 400, kHop = 160, kBins = 257, kHalf = 256;
constexpr int kMaxSegsSmem = 2048;       // segments whose frame prefix is staged in shared memory

constexpr int kMelGroups = kNMels / 32;
struct FrontTablesHost {
  std::vector<float> window;       // [400]
  std::vector<int> mel_lo, mel_cnt, mel_off;  // [128]
  std::vector<float> mel_w;        // packed nonzero weights
  // transposed copy for logmel_reg_kernel: group g = filters 32g .. 32g+31, depth[g] = longest triangle of the group,
  // weight q of filter 32g + lane at mel_wt[off_t[g] + 32 q + lane] (zero beyond the filter's own count)
  int depth[kMelGroups], off_t[kMelGroups];
  std::vector<float> mel_wt;
};

// Same f32 arithmetic as lib.rs:174-223 (the tables are data, built once on the host).
static FrontTablesHost build_tables() {
  FrontTablesHost t;
  t.window.resize(kWin);
  for (int i = 0; i < kWin; ++i)
    t.window[i] = 0.5f * (1.0f - cosf(2.0f * 3.14159265358979323846f * (float)i / (float)(kWin - 1)));
  auto hz_to_mel = [](float hz) { return 2595.0f * log10f(1.0f + hz / 700.0f); };
  auto mel_to_hz = [](float mel) { return 700.0f * (powf(10.0f, mel / 2595.0f) - 1.0f); };
  const float min_mel = hz_to_mel(0.0f), max_mel = hz_to_mel(8000.0f);
  float pts[kNMels + 2];
  for (int i = 0; i < kNMels + 2; ++i) pts[i] = mel_to_hz(min_mel + (max_mel - min_mel) * ((float)i / (float)(kNMels + 1)));
  t.mel_lo.resize(kNMels); t.mel_cnt.resize(kNMels); t.mel_off.resize(kNMels);
  for (int m = 0; m < kNMels; ++m) {
    const float left = pts[m], center = pts[m + 1], right = pts[m + 2];
    int lo = -1, hi = -1;
    std::vector<float> row(kBins, 0.0f);
    for (int i = 0; i < kBins; ++i) {
      const float freq = (float)i * 16000.0f / (float)kNfft;
      float w = 0.0f;
      if (freq > left && freq < center) w = (freq - left) / (center - left);
      else if (freq >= center && freq < right) w = (right - freq) / (right - center);
      row[i] = w;
      if (w != 0.0f) { if (lo < 0) lo = i; hi = i; }
    }
    t.mel_off[m] = (int)t.mel_w.size();
    t.mel_lo[m] = lo < 0 ? 0 : lo;
    t.mel_cnt[m] = lo < 0 ? 0 : hi - lo + 1;
    for (int i = 0; i < t.mel_cnt[m]; ++i) t.mel_w.push_back(row[t.mel_lo[m] + i]);
  }
  for (int g = 0; g < kMelGroups; ++g) {
    int d = 0;
    for (int l = 0; l < 32; ++l) d = t.mel_cnt[32 * g + l] > d ? t.mel_cnt[32 * g + l] : d;
    t.depth[g] = d;
    t.off_t[g] = (int)t.mel_wt.size();
    for (int q = 0; q < d; ++q)
      for (int l = 0; l < 32; ++l) {
        const int m = 32 * g + l;
        t.mel_wt.push_back(q < t.mel_cnt[m] ? t.mel_w[t.mel_off[m] + q] : 0.0f);
      }
  }
  return t;
}

struct FrontTablesDev {
  float* window; int* mel_lo; int* mel_cnt; int* mel_off; float* mel_w; int n_w;
  float* mel_wt; int n_wt; int depth[kMelGroups]; int off_t[kMelGroups];
};

Frontend::Frontend() {
  FrontTablesHost h = build_tables();
  auto* d = new FrontTablesDev();
  PKB_CUDA(cudaMalloc(&d->window, kWin * 4));
  PKB_CUDA(cudaMalloc(&d->mel_lo, kNMels * 4));
  PKB_CUDA(cudaMalloc(&d->mel_cnt, kNMels * 4));
  PKB_CUDA(cudaMalloc(&d->mel_off, kNMels * 4));
  PKB_CUDA(cudaMalloc(&d->mel_w, h.mel_w.size() * 4));
  d->n_w = (int)h.mel_w.size();
  PKB_CHECK(d->n_w <= 768, "mel filterbank has more non-zeros than the shared-memory table");
  PKB_CUDA(cudaMemcpy(d->window, h.window.data(), kWin * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(d->mel_lo, h.mel_lo.data(), kNMels * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(d->mel_cnt, h.mel_cnt.data(), kNMels * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(d->mel_off, h.mel_off.data(), kNMels * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(d->mel_w, h.mel_w.data(), h.mel_w.size() * 4, cudaMemcpyHostToDevice));
  d->n_wt = (int)h.mel_wt.size();
  PKB_CHECK(d->n_wt <= 768, "transposed mel filterbank does not fit the shared-memory table");
  for (int g = 0; g < kMelGroups; ++g) {
    d->depth[g] = h.depth[g]; d->off_t[g] = h.off_t[g];
    // the zero-padded tail of a short triangle still reads power[lo + q]: it must stay inside the padded power buffer
    for (int l = 0; l < 32; ++l) PKB_CHECK(h.mel_lo[32 * g + l] + h.depth[g] <= kBins + 15, "mel triangle reads past the power buffer");
  }
  PKB_CUDA(cudaMalloc(&d->mel_wt, h.mel_wt.size() * 4));
  PKB_CUDA(cudaMemcpy(d->mel_wt, h.mel_wt.data(), h.mel_wt.size() * 4, cudaMemcpyHostToDevice));
  tables_ = d;
}

Frontend::~Frontend() {
  auto* d = static_cast<FrontTablesDev*>(tables_);
  if (!d) return;
  cudaFree(d->window); cudaFree(d->mel_lo); cudaFree(d->mel_cnt); cudaFree(d->mel_off); cudaFree(d->mel_w); cudaFree(d->mel_wt);
  delete d;
}

// ------------------------------------------------------------------------------------------------
// shared memory per CTA: tables (window 400, tw256 256x2, tw512 257x2, mel meta 3x128, mel weights <=640)
// + per-warp 2 x 256 complex ping-pong + 257 power values.
struct __align__(16) WarpBuf {
  float2 a[kHalf];
  float2 b[kHalf];
};

__device__ __forceinline__ float2 cmul(float2 x, float2 w) { return make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x); }

__global__ void __launch_bounds__(kWarpsPerCta * 32)
logmel_kernel(const float* __restrict__ audio, const FrontSegment* __restrict__ segs, const int* __restrict__ frame_prefix,
              int n_segs, int total_frames, FrontTablesDev tb, float* __restrict__ out, const float* __restrict__ stats) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_window = reinterpret_cast<float*>(smem_raw);            // 400
  float2* s_tw256 = reinterpret_cast<float2*>(s_window + kWin);    // 256: e^{-2 pi i k/256}
  float2* s_tw512 = s_tw256 + kHalf;                                // 257: e^{-2 pi i k/512}
  int* s_lo = reinterpret_cast<int*>(s_tw512 + kBins + 1);          // 128
  int* s_cnt = s_lo + kNMels;
  int* s_off = s_cnt + kNMels;
  float* s_w = reinterpret_cast<float*>(s_off + kNMels);            // n_w (<= 768)
  WarpBuf* s_warp = reinterpret_cast<WarpBuf*>(s_w + 768);
  int* s_prefix = reinterpret_cast<int*>(s_warp + kWarpsPerCta);   // [kMaxSegsSmem + 1] frame prefix (binary-searched per frame)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kWin; i += blockDim.x) s_window[i] = tb.window[i];
  for (int i = tid; i < kHalf; i += blockDim.x) {
    float s, c;
    sincospif(-2.0f * (float)i / 256.0f, &s, &c);
    s_tw256[i] = make_float2(c, s);
  }
  for (int i = tid; i < kBins; i += blockDim.x) {
    float s, c;
    sincospif(-2.0f * (float)i / 512.0f, &s, &c);
    s_tw512[i] = make_float2(c, s);
  }
  for (int i = tid; i < kNMels; i += blockDim.x) { s_lo[i] = tb.mel_lo[i]; s_cnt[i] = tb.mel_cnt[i]; s_off[i] = tb.mel_off[i]; }
  for (int i = tid; i < tb.n_w; i += blockDim.x) s_w[i] = tb.mel_w[i];
  const bool prefix_in_smem = n_segs <= kMaxSegsSmem;
  if (prefix_in_smem)
    for (int i = tid; i <= n_segs; i += blockDim.x) s_prefix[i] = frame_prefix[i];
  const int* prefix = prefix_in_smem ? s_prefix : frame_prefix;      // 10 dependent global loads per frame otherwise
  __syncthreads();

  WarpBuf& wb = s_warp[warp];
  const int warps_total = gridDim.x * kWarpsPerCta;
  for (int gf = blockIdx.x * kWarpsPerCta + warp; gf < total_frames; gf += warps_total) {
    // locate the segment of global frame gf (frame_prefix[s] = first global frame of segment s)
    int lo = 0, hi = n_segs - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= gf) lo = mid; else hi = mid - 1;
    }
    const FrontSegment sg = segs[lo];
    const int t = gf - prefix[lo];
    const float* x = audio + sg.audio_off + (size_t)t * kHop;

    // windowed frame -> packed complex z[n] = x[2n] + i x[2n+1]; samples >= 400 are zero
#pragma unroll
    for (int i = 0; i < kHalf / 32; ++i) {
      const int n = lane + 32 * i;
      float2 v = make_float2(0.f, 0.f);
      if (2 * n < kWin) {
        const float2 xv = *reinterpret_cast<const float2*>(x + 2 * n);   // audio_off is even by construction
        v.x = xv.x * s_window[2 * n];
        v.y = xv.y * s_window[2 * n + 1];
      }
      wb.a[n] = v;
    }
    __syncwarp();

    // radix-4 Stockham: Ns = 1, 4, 16, 64
    float2* src = wb.a;
    float2* dst = wb.b;
#pragma unroll
    for (int stage = 0; stage < 4; ++stage) {
      const int Ns = 1 << (2 * stage);
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const int j = lane + 32 * rep;                  // butterfly index 0..63
        const int m = j & (Ns - 1);
        float2 v0 = src[j], v1 = src[j + 64], v2 = src[j + 128], v3 = src[j + 192];
        const int tw = m * (64 / Ns);                   // angle = -2 pi m / (4 Ns) = -2 pi tw / 256
        v1 = cmul(v1, s_tw256[tw]);
        v2 = cmul(v2, s_tw256[2 * tw]);
        v3 = cmul(v3, s_tw256[3 * tw]);
        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
        const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
        const float2 a3 = make_float2(d13.y, -d13.x);   // -i * (v1 - v3)
        const int base = (j / Ns) * (Ns * 4) + m;
        dst[base] = make_float2(a0.x + a2.x, a0.y + a2.y);
        dst[base + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
        dst[base + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
        dst[base + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
      }
      __syncwarp();
      float2* tmp = src; src = dst; dst = tmp;
    }
    // src now holds Z[0..255].  Power spectrum of the 512-point real FFT into dst (as floats).
    float* pw = reinterpret_cast<float*>(dst);
    for (int k = lane; k <= kHalf; k += 32) {
      const float2 zk = src[k & (kHalf - 1)];
      const float2 zn = src[(kHalf - k) & (kHalf - 1)];
      const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));        // (Z[k] + conj Z[N-k]) / 2
      const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));       // (Z[k] - conj Z[N-k]) / (2i)
      const float2 ow = cmul(o, s_tw512[k]);
      const float re = e.x + ow.x, im = e.y + ow.y;
      pw[k] = re * re + im * im;
    }
    __syncwarp();
    // sparse mel + log (+ optional per-feature normalisation with caller-provided stats)
    const int orow = sg.ring_cap > 0 ? (sg.frame0 + t) % sg.ring_cap : t;
    float* o = out + sg.out_off + (size_t)orow * sg.out_stride;
#pragma unroll
    for (int i = 0; i < kNMels / 32; ++i) {
      const int m = lane + 32 * i;
      const int b0 = s_lo[m], cnt = s_cnt[m], off = s_off[m];
      float e = 0.0f;
      for (int q = 0; q < cnt; ++q) e += pw[b0 + q] * s_w[off + q];
      float v = logf(e + 1e-5f);
      if (sg.norm_off >= 0) v = (v - stats[sg.norm_off + m]) / stats[sg.norm_off + kNMels + m];
      o[m] = v;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Register-resident version.  Index algebra (n = input index, k = output bin of the 256-point transform, W_N = e^{-2 pi i / N}):
//   n = n0 + 32 n1 (lane n0, register n1), k = k1 + 8 k2:   Z[k] = sum_n0 W_32^{n0 k2} . W_256^{n0 k1} . (sum_n1 W_8^{n1 k1} z[n])
//   n0 = n00 + 4 n01, k2 = k20 + 8 k21:   W_32^{n0 k2} = W_8^{n01 k20} . W_32^{n00 k20} . W_4^{n00 k21}
// pass 1: lane n0 -- radix-8 over n1, twiddle W_256^{n0 k1};   transpose 1: lane k1 + 8 n00 gathers n01 = 0..7
// pass 2: radix-8 over n01, twiddle W_32^{n00 k20};            transpose 2: lane k1 + 8 (k20 & 3) gathers n00 = 0..3 for k20, k20 + 4
// pass 3: two radix-4 over n00 -> the lane holds Z[lane + 32 i] in register i = (k20 >> 2) + 2 k21
// Transpose pitches 34 / 40 (float2 units) make every 64-bit access of a half-warp hit 16 distinct 8-byte banks.
constexpr int kT1Pitch = 34, kT2Pitch = 40;
struct __align__(16) WarpBufR {
  float2 t[8 * kT2Pitch];
  float pw[kBins + 15];       // power spectrum; the tail stays zero (read by the zero-weight padding of short triangles)
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// forward 4-point DFT: X1 = (x0 - x2) - i (x1 - x3), X3 = (x0 - x2) + i (x1 - x3)
__device__ __forceinline__ void dft4(float2 x0, float2 x1, float2 x2, float2 x3, float2& X0, float2& X1, float2& X2, float2& X3) {
  const float2 s0 = cadd(x0, x2), d0 = csub(x0, x2), s1 = cadd(x1, x3), d1 = csub(x1, x3);
  X0 = cadd(s0, s1);
  X2 = csub(s0, s1);
  X1 = make_float2(d0.x + d1.y, d0.y - d1.x);
  X3 = make_float2(d0.x - d1.y, d0.y + d1.x);
}
// forward 8-point DFT in place, natural order in and out (decimation in time: even / odd 4-point transforms, then W_8^k)
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  float2 e0, e1, e2, e3, o0, o1, o2, o3;
  dft4(a[0], a[2], a[4], a[6], e0, e1, e2, e3);
  dft4(a[1], a[3], a[5], a[7], o0, o1, o2, o3);
  const float r = 0.70710678118654752440f;
  const float2 t0 = o0;
  const float2 t1 = make_float2(r * (o1.x + o1.y), r * (o1.y - o1.x));       // o1 . (1 - i) / sqrt 2
  const float2 t2 = make_float2(o2.y, -o2.x);                                // o2 . (-i)
  const float2 t3 = make_float2(r * (o3.y - o3.x), -r * (o3.x + o3.y));      // o3 . (-1 - i) / sqrt 2
  a[0] = cadd(e0, t0); a[4] = csub(e0, t0);
  a[1] = cadd(e1, t1); a[5] = csub(e1, t1);
  a[2] = cadd(e2, t2); a[6] = csub(e2, t2);
  a[3] = cadd(e3, t3); a[7] = csub(e3, t3);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
logmel_reg_kernel(const float* __restrict__ audio, const FrontSegment* __restrict__ segs, const int* __restrict__ frame_prefix,
                  int n_segs, int total_frames, FrontTablesDev tb, float* __restrict__ out, const float* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_wt = reinterpret_cast<float*>(smem_raw);                          // 768: transposed mel weights
  WarpBufR* s_warp = reinterpret_cast<WarpBufR*>(s_wt + 768);
  int* s_prefix = reinterpret_cast<int*>(s_warp + kWarpsPerCta);            // [kMaxSegsSmem + 1]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tables are constants of the library: staged before the dependency wait
  for (int i = tid; i < tb.n_wt; i += blockDim.x) s_wt[i] = tb.mel_wt[i];
  WarpBufR& wb = s_warp[warp];
  for (int i = lane; i < kBins + 15; i += 32) wb.pw[i] = 0.0f;
  // per-lane loop invariants: window samples of the lane's 7 non-zero points, twiddles of the three passes, mel triangle starts
  float2 win[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int n = lane + 32 * i;
    win[i] = 2 * n < kWin ? make_float2(tb.window[2 * n], tb.window[2 * n + 1]) : make_float2(0.f, 0.f);
  }
  float2 tw1[7], tw2[7];
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    float sn, cs;
    sincospif(-2.0f * (float)(lane * k) / 256.0f, &sn, &cs);
    tw1[k - 1] = make_float2(cs, sn);
    sincospif(-2.0f * (float)((lane >> 3) * k) / 32.0f, &sn, &cs);
    tw2[k - 1] = make_float2(cs, sn);
  }
  float2 twl;
  { float sn, cs; sincospif(-2.0f * (float)lane / 512.0f, &sn, &cs); twl = make_float2(cs, sn); }
  int mel_b0[kMelGroups];
#pragma unroll
  for (int g = 0; g < kMelGroups; ++g) mel_b0[g] = tb.mel_lo[32 * g + lane];

  pdl_enter();
  const bool prefix_in_smem = n_segs <= kMaxSegsSmem;
  if (prefix_in_smem)
    for (int i = tid; i <= n_segs; i += blockDim.x) s_prefix[i] = frame_prefix[i];
  const int* prefix = prefix_in_smem ? s_prefix : frame_prefix;
  __syncthreads();

  const int warps_total = gridDim.x * kWarpsPerCta;
  for (int gf = blockIdx.x * kWarpsPerCta + warp; gf < total_frames; gf += warps_total) {
    int lo = 0, hi = n_segs - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= gf) lo = mid; else hi = mid - 1;
    }
    const FrontSegment sg = segs[lo];
    const int t = gf - prefix[lo];
    const float* x = audio + sg.audio_off + (size_t)t * kHop;

    // windowed frame straight into registers: z[n] = x[2n] + i x[2n+1], n = lane + 32 i; samples >= 400 are zero
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = lane + 32 * i;
      a[i] = make_float2(0.f, 0.f);
      if (i < 7 && 2 * n < kWin) {
        const float2 xv = *reinterpret_cast<const float2*>(x + 2 * n);   // audio_off is even by construction
        a[i] = make_float2(xv.x * win[i].x, xv.y * win[i].y);
      }
    }
    // pass 1
    dft8(a);
#pragma unroll
    for (int k = 1; k < 8; ++k) a[k] = cmul(a[k], tw1[k - 1]);
#pragma unroll
    for (int k = 0; k < 8; ++k) wb.t[kT1Pitch * k + lane] = a[k];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = wb.t[kT1Pitch * (lane & 7) + (lane >> 3) + 4 * j];
    __syncwarp();
    // pass 2
    dft8(a);
#pragma unroll
    for (int k = 1; k < 8; ++k) a[k] = cmul(a[k], tw2[k - 1]);
#pragma unroll
    for (int k = 0; k < 8; ++k) wb.t[kT2Pitch * k + lane] = a[k];
    __syncwarp();
    // pass 3: z[j + 2 k21] = Z[lane + 32 (j + 2 k21)]
    float2 z[8];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2* src = wb.t + kT2Pitch * ((lane >> 3) + 4 * j) + (lane & 7);
      dft4(src[0], src[8], src[16], src[24], z[j], z[j + 2], z[j + 4], z[j + 6]);
    }
    // even/odd split: bin k = lane + 32 i pairs with Z[256 - k] = register 7 - i of lane 32 - lane (lane 0: its own register 8 - i)
    const int partner = (32 - lane) & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float2 zn;
      zn.x = __shfl_sync(0xffffffffu, z[7 - i].x, partner);
      zn.y = __shfl_sync(0xffffffffu, z[7 - i].y, partner);
      if (lane == 0) zn = z[(8 - i) & 7];
      const float2 zk = z[i];
      const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));        // (Z[k] + conj Z[N-k]) / 2
      const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));       // (Z[k] - conj Z[N-k]) / (2i)
      // W_512^k = W_512^lane . W_16^i
      constexpr float c16[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
      constexpr float s16[8] = {0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                                -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
      const float2 ow = cmul(o, cmul(twl, make_float2(c16[i], s16[i])));
      const float re = e.x + ow.x, im = e.y + ow.y;
      wb.pw[lane + 32 * i] = re * re + im * im;
    }
    if (lane == 0) { const float d = z[0].x - z[0].y; wb.pw[kHalf] = d * d; }      // bin 256: Re Z[0] - Im Z[0]
    __syncwarp();
    // sparse mel + log (+ optional per-feature normalisation with caller-provided stats)
    const int orow = sg.ring_cap > 0 ? (sg.frame0 + t) % sg.ring_cap : t;
    float* o = out + sg.out_off + (size_t)orow * sg.out_stride;
#pragma unroll
    for (int g = 0; g < kMelGroups; ++g) {
      const float* pw = wb.pw + mel_b0[g];
      const float* w = s_wt + tb.off_t[g] + lane;
      float e = 0.0f;
      for (int q = 0; q < tb.depth[g]; ++q) e += pw[q] * w[32 * q];
      float v = logf(e + 1e-5f);
      const int m = lane + 32 * g;
      if (sg.norm_off >= 0) v = (v - stats[sg.norm_off + m]) / stats[sg.norm_off + kNMels + m];
      o[m] = v;
    }
    __syncwarp();
  }
}

static size_t logmel_reg_smem_bytes() { return 768 * 4 + sizeof(WarpBufR) * kWarpsPerCta + (kMaxSegsSmem + 1) * 4; }
static size_t logmel_smem_bytes() {
  return (kWin + 2 * kHalf + 2 * (kBins + 1) + 3 * kNMels + 768) * 4 + sizeof(WarpBuf) * kWarpsPerCta + (kMaxSegsSmem + 1) * 4;
}

void Frontend::logmel(const float* d_audio, const FrontSegment* d_segs, const int* d_frame_prefix, int n_segs,
                      int total_frames, float* d_out, const float* d_stats, int sm_count, cudaStream_t st) {
  if (total_frames <= 0) return;
  auto* d = static_cast<FrontTablesDev*>(tables_);
  static const int use_reg = [] { const char* v = getenv("PARAKEET_B200_LOGMEL"); return (v && v[0] == '0') ? 0 : 1; }();
  if (use_reg) {
    const size_t smem = logmel_reg_smem_bytes();
    static bool attr_reg = false;
    if (!attr_reg) {
      PKB_CUDA(cudaFuncSetAttribute(logmel_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_reg = true;
    }
    int ctas = (total_frames + kWarpsPerCta - 1) / kWarpsPerCta;
    const int cap = sm_count * 2;   // two resident 8-warp CTAs per SM, grid-stride over frames
    if (ctas > cap) ctas = cap;
    launch_k(logmel_reg_kernel, dim3(ctas), dim3(kWarpsPerCta * 32), smem, st, d_audio, d_segs, d_frame_prefix, n_segs, total_frames, *d, d_out,
             d_stats);
    PKB_CUDA(cudaGetLastError());
    return;
  }
  const size_t smem = logmel_smem_bytes();
  static bool attr_set = false;
  if (!attr_set) {
    PKB_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int ctas = (total_frames + kWarpsPerCta - 1) / kWarpsPerCta;
  const int cap = sm_count * 4;   // persistent-style: <= 4 CTAs per SM, grid-stride over frames
  if (ctas > cap) ctas = cap;
  launch_k(logmel_kernel, dim3(ctas), dim3(kWarpsPerCta * 32), smem, st, d_audio, d_segs, d_frame_prefix, n_segs, total_frames, *d, d_out, d_stats);
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Per-feature statistics over one utterance (lib.rs:127-159): mean over T, std = sqrt(sum (x-mu)^2/(T-1)) + 1e-5.
// One CTA per utterance, thread == mel bin, frames summed SEQUENTIALLY in f32 exactly like the reference loop: the
// empty mel filter 0 makes feature 0 a constant whose normalised value is pure summation-order noise divided by the
// 1e-5 std floor, so only the same order reproduces the reference there.  Rows are 512-byte coalesced reads; the
// dependent add chain costs ~4 cycles per frame (1.5 ms for a one-hour utterance), utterances run on separate SMs.
__global__ void __launch_bounds__(128)
feature_stats_kernel(const float* __restrict__ feat, const FrontSegment* __restrict__ segs, const int* __restrict__ frames,
                     float* __restrict__ stats /* [n_segs][2][128] */) {
  const int s = blockIdx.x, m = threadIdx.x;
  const int T = frames[s];
  const float* f = feat + segs[s].out_off + m;
  const size_t stride = (size_t)segs[s].out_stride;
  float acc = 0.0f;
  int t = 0;
  for (; t + 8 <= T; t += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f[(size_t)(t + i) * stride];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i];
  }
  for (; t < T; ++t) acc += f[(size_t)t * stride];
  const float mu = T > 0 ? acc / (float)T : 0.0f;
  acc = 0.0f;
  t = 0;
  for (; t + 8 <= T; t += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f[(size_t)(t + i) * stride];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mu; acc = __fadd_rn(acc, __fmul_rn(d, d)); }
  }
  for (; t < T; ++t) { const float d = f[(size_t)t * stride] - mu; acc = __fadd_rn(acc, __fmul_rn(d, d)); }
  const float denom = T > 1 ? (float)(T - 1) : 1.0f;
  stats[(size_t)s * 256 + m] = mu;
  stats[(size_t)s * 256 + 128 + m] = T > 0 ? sqrtf(acc / denom) + 1e-5f : 0.0f;
}

__global__ void __launch_bounds__(256)
feature_norm_kernel(float* __restrict__ feat, const FrontSegment* __restrict__ segs, const int* __restrict__ frames,
                    const float* __restrict__ stats, int max_frames) {
  const int s = blockIdx.y;
  const int T = frames[s];
  float* f = feat + segs[s].out_off;
  const int stride = segs[s].out_stride;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over T*128/4 float4s
  const size_t t = idx / 32;
  const int m4 = (int)(idx % 32) * 4;
  if (t >= (size_t)T) return;
  float4 v = *reinterpret_cast<float4*>(f + t * stride + m4);
  const float4 mu = *reinterpret_cast<const float4*>(stats + (size_t)s * 256 + m4);
  const float4 sd = *reinterpret_cast<const float4*>(stats + (size_t)s * 256 + 128 + m4);
  v.x = (v.x - mu.x) / sd.x; v.y = (v.y - mu.y) / sd.y; v.z = (v.z - mu.z) / sd.z; v.w = (v.w - mu.w) / sd.w;
  *reinterpret_cast<float4*>(f + t * stride + m4) = v;
  (void)max_frames;
}

// One CTA per segment, thread == mel bin: sequential over the segment's new frames (24 per streaming step).
__global__ void __launch_bounds__(128)
running_norm_kernel(float* __restrict__ out, const FrontSegment* __restrict__ segs, const int* __restrict__ frame_prefix,
                    float* __restrict__ state) {
  const FrontSegment sg = segs[blockIdx.x];
  if (sg.run_state == 0) return;
  const int m = threadIdx.x;
  float* st = state + (sg.run_state - 1);
  const int T = frame_prefix[blockIdx.x + 1] - frame_prefix[blockIdx.x];
  float n = st[0], mean = st[1 + m], m2 = st[1 + kNMels + m];
  __syncthreads();      // every thread has read the shared count before thread 0 updates it
  for (int t = 0; t < T; ++t) {
    const int orow = sg.ring_cap > 0 ? (sg.frame0 + t) % sg.ring_cap : t;
    float* o = out + sg.out_off + (size_t)orow * sg.out_stride + m;
    const float x = *o;
    n += 1.0f;
    const float delta = x - mean;
    mean += delta / n;
    m2 = fmaf(delta, x - mean, m2);
    *o = n > 1.5f ? (x - mean) / (sqrtf(m2 / (n - 1.0f)) + 1e-5f) : 0.0f;
  }
  st[1 + m] = mean;
  st[1 + kNMels + m] = m2;
  if (m == 0) st[0] = n;
}

void Frontend::running_norm(float* d_out, const FrontSegment* d_segs, const int* d_frame_prefix, int n_segs, float* d_state,
                            cudaStream_t st) {
  if (n_segs <= 0) return;
  running_norm_kernel<<<n_segs, kNMels, 0, st>>>(d_out, d_segs, d_frame_prefix, d_state);
  PKB_CUDA(cudaGetLastError());
}

void Frontend::per_feature_stats(const float* d_feat, const FrontSegment* d_segs, const int* d_frames, int n_segs,
                                 float* d_stats, cudaStream_t st) {
  if (n_segs <= 0) return;
  feature_stats_kernel<<<n_segs, kNMels, 0, st>>>(d_feat, d_segs, d_frames, d_stats);
  PKB_CUDA(cudaGetLastError());
}

void Frontend::apply_norm(float* d_feat, const FrontSegment* d_segs, const int* d_frames, int n_segs, int max_frames,
                          const float* d_stats, cudaStream_t st) {
  if (n_segs <= 0 || max_frames <= 0) return;
  const int blocks = (int)(((size_t)max_frames * 32 + 255) / 256);
  feature_norm_kernel<<<dim3(blocks, n_segs), 256, 0, st>>>(d_feat, d_segs, d_frames, d_stats, max_frames);
  PKB_CUDA(cudaGetLastError());
}

// [128,T] bins-major (the C-ABI layout, parakeet_trt.cpp:2001-2004) -> frames-major rows of the feature ring.
__global__ void __launch_bounds__(256)
bins_to_frames_kernel(const float* __restrict__ src, int T, float* __restrict__ ring, int ring_cap, int frame0) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + tx;
    tile[r][tx] = t < T ? src[(size_t)(m0 + r) * T + t] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r;
    if (t < T) ring[(size_t)((frame0 + t) % ring_cap) * kNMels + m0 + tx] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(256)
bins_to_frames_batch_kernel(const float* __restrict__ stage, const FeatPush* __restrict__ push, float* __restrict__ rings, int ring_cap) {
  __shared__ float tile[32][33];
  const FeatPush p = push[blockIdx.z];
  const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  if (t0 >= p.T) return;
  const float* src = stage + p.src_off;
  float* ring = rings + p.ring_off;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + tx;
    tile[r][tx] = t < p.T ? src[(size_t)(m0 + r) * p.T + t] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r;
    if (t < p.T) ring[(size_t)((p.frame0 + t) % ring_cap) * kNMels + m0 + tx] = tile[tx][r];
  }
}

void Frontend::bins_to_frames_batch(const float* d_stage, const FeatPush* d_push, int n, int max_T, float* d_rings, int ring_cap,
                                    cudaStream_t st) {
  if (n <= 0 || max_T <= 0) return;
  for (int i0 = 0; i0 < n; i0 += 65535) {      // grid.z limit
    const int cnt = n - i0 < 65535 ? n - i0 : 65535;
    bins_to_frames_batch_kernel<<<dim3((max_T + 31) / 32, kNMels / 32, cnt), 256, 0, st>>>(d_stage, d_push + i0, d_rings, ring_cap);
  }
  PKB_CUDA(cudaGetLastError());
}

void Frontend::bins_to_frames(const float* d_src, int T, float* d_ring, int ring_cap, int frame0, cudaStream_t st) {
  if (T <= 0) return;
  bins_to_frames_kernel<<<dim3((T + 31) / 32, kNMels / 32), 256, 0, st>>>(d_src, T, d_ring, ring_cap, frame0);
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
