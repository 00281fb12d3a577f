// frontend.cu -- log-mel frontend on sm_100a (piece 1 of the hot path).
//
// Replaces the CPU extractor /root/reference/rust/features/src/lib.rs:66-120 (framing, symmetric Hann-400,
// tail zero-pad to 512, real FFT, |X|^2, 128 un-normalised HTK-mel triangles, ln(E+1e-5)) and the
// per-feature normalisation :127-172.  HBM-bound by design: one pass reads 16000 samples and writes
// 100x128 features per audio second (115 200 B); everything in between lives in shared memory / registers.
//
// Mapping: one warp per frame.  The 512-point real FFT is done as a 256-point complex FFT
// (z[n] = x[2n] + i x[2n+1]) with a radix-4 Stockham autosort in the warp's shared-memory slice
// (4 stages, 2 radix-4 butterflies per lane per stage), followed by the even/odd split.  The mel
// stage is sparse: each triangle touches a contiguous bin range, summed in bin order.
#include "frontend.h"

#include <math.h>

#include <vector>

namespace pkb {

constexpr int kWarpsPerCta = 8;
constexpr int kNfft = 512, kWin = 400, kHop = 160, kBins = 257, kHalf = 256;
constexpr int kMaxSegsSmem = 2048;       // segments whose frame prefix is staged in shared memory

struct FrontTablesHost {
  std::vector<float> window;       // [400]
  std::vector<int> mel_lo, mel_cnt, mel_off;  // [128]
  std::vector<float> mel_w;        // packed nonzero weights
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
  return t;
}

struct FrontTablesDev {
  float* window; int* mel_lo; int* mel_cnt; int* mel_off; float* mel_w; int n_w;
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
  tables_ = d;
}

Frontend::~Frontend() {
  auto* d = static_cast<FrontTablesDev*>(tables_);
  if (!d) return;
  cudaFree(d->window); cudaFree(d->mel_lo); cudaFree(d->mel_cnt); cudaFree(d->mel_off); cudaFree(d->mel_w);
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

static size_t logmel_smem_bytes() {
  return (kWin + 2 * kHalf + 2 * (kBins + 1) + 3 * kNMels + 768) * 4 + sizeof(WarpBuf) * kWarpsPerCta + (kMaxSegsSmem + 1) * 4;
}

void Frontend::logmel(const float* d_audio, const FrontSegment* d_segs, const int* d_frame_prefix, int n_segs,
                      int total_frames, float* d_out, const float* d_stats, int sm_count, cudaStream_t st) {
  if (total_frames <= 0) return;
  auto* d = static_cast<FrontTablesDev*>(tables_);
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
