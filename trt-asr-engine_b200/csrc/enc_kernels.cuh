// enc_kernels.cuh -- non-GEMM kernels of the cache-aware FastConformer chunk (piece 2) and the per-stream
// cache carry-over (piece 4).  See enc_kernels.cu for the mapping of each kernel.
#pragma once
#include "common.cuh"

namespace pkb {

// Per-step batch descriptor.  One "entry" = one stream advancing by one chunk in this batched step.
// All arrays live in one device int buffer, written by a single H2D copy per step (see Engine::upload_batch).
struct BatchDev {
  int B = 0;          // entries
  int M = 0;          // packed encoder rows = sum Tq
  int sumT2 = 0, sumT3 = 0;
  int max_Tq = 0;
  int max_tenc = 0;   // most encoder frames any entry emits (3 streaming, Tq offline)
  const int* slot = nullptr;    // [B] state slot of the stream
  const int* T = nullptr;       // [B] input feature frames of this chunk
  const int* f0 = nullptr;      // [B] absolute index of the chunk's first frame in the stream's feature ring
  const int* T1 = nullptr;      // [B] frames after conv0 ((T-1)/2+1)
  const int* T2 = nullptr;      // [B] after dw1
  const int* T3 = nullptr;      // [B] after dw2 (pre-encode tokens)
  const int* Tq = nullptr;      // [B] T3 - drop_extra_pre_encoded
  const int* qlen = nullptr;    // [B] valid new tokens (length after subsampling - 2, clamped to [0,Tq])
  const int* len = nullptr;     // [B] cache_last_channel_len on entry
  const int* head = nullptr;    // [B] physical ring index of logical cache position 0
  const int* drop = nullptr;    // [B] pre-encode tokens dropped at the front (2 streaming = drop_extra_pre_encoded, 0 offline)
  const int* offline = nullptr; // [B] 1: offline (full-context) encode of this push: no caches read or carried over
  const int* off2 = nullptr;    // [B+1] prefix sums of T2
  const int* off3 = nullptr;    // [B+1] prefix sums of T3
  const int* row_off = nullptr; // [B+1] prefix sums of Tq
  int* row_entry = nullptr;     // [M]
  int* row_pos = nullptr;       // [M]
  int* rowmap3 = nullptr;       // [sumT3] -> packed row or -1 (dropped pre-encode token)
};

struct SubsampleWeights {
  const float* w0; const float* b0;   // conv.0  [256,1,3,3], [256]
  const float* w2; const float* b2;   // conv.2  depthwise
  const float* w5; const float* b5;   // conv.5  depthwise
  // the three 3x3 filters transposed to [9 taps][256 channels]: lanes that own adjacent channels read adjacent words (with the
  // [256][9] layout every weight request of a warp touched up to 32 sectors: 163 M load sectors per stage-1 launch at 1024 streams)
  const float* w0t; const float* w2t; const float* w5t;
};

struct ActOut {             // destination of a GEMM-A operand (bf16 hi plane [+ lo plane])
  __nv_bfloat16* ptr;
  int lda;
  long long lo_off;
};

void launch_build_rows(const BatchDev& b, cudaStream_t st);

// conv0(1->256,3x3,s2,p1)+ReLU fused with depthwise conv.2 (3x3,s2,p1): feature ring -> A1 [sumT2*32, 256]
void launch_subsample_stage1(const BatchDev& b, const float* feat_ring, int ring_cap, const SubsampleWeights& w, ActOut a1,
                             cudaStream_t st);
// depthwise conv.5 over y1 [sumT2*32, 256] (channels-last; bf16 hi plane [+ lo plane in precise mode], as written by the pointwise
// GEMM's EPI_BIAS_RELU_ACT epilogue) -> A2 [sumT3*16, 256]
void launch_subsample_stage2(const BatchDev& b, ActOut y1, const SubsampleWeights& w, ActOut a2, cudaStream_t st);

// LayerNorm over rows of x f32 [M,1024] (one warp per row, eps 1e-5).
//  write_x == 0:  A <- LN1(x)
//  write_x == 1:  x <- LN1(x) in place; A <- (g2 ? LN2(x) : x) if A.ptr   (norm_out fused with the next layer's
//                 norm_feed_forward1; after the last layer A is the joint's encoder-projection operand)
//  acache != null: LN1(x) rows are also stored in the contract cache ring (cache_last_channel, pre-projection)
// deferred residual (split-K GEMM partial sums, see gemm.h EPI_PARTIAL_F32): x += scale * sum_s part[s]
struct LnResidual { const float* part; int splits; long long split_stride; float scale; int bf16; };   // bf16: partials are bf16 elements
struct AcacheOut { void* ring; int is_f32; const int* row_entry; const int* row_pos; const int* entry_slot; const int* entry_head; };
void launch_layernorm(float* x, int M, const float* g1, const float* b1, const float* g2, const float* b2, int write_x, ActOut a,
                      const AcacheOut* ac, cudaStream_t st, const LnResidual* res = nullptr);
// LayerNorm folded into the A operand of the CUDA-core GEMM (gemm_simt_ln, one-stream latency path): A = LN(x) with gamma / beta
struct LnFuse { const float* x; const float* gamma; const float* beta; AcacheOut ac; int has_ac; };
#ifdef __CUDACC__
// One LayerNorm row held by a warp: lane owns columns i*128 + lane*4 .. +3 (i = 0..7) in v[4i .. 4i+3]; two-pass mean / variance,
// eps 1e-5.  Every LayerNorm in the library goes through this function (same summation order everywhere).  The elementwise passes
// use the packed f32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest results per issue slot) -- the
// row kernels are bound by instruction issue, not by memory (profiles/r02_ln_stream...).
__device__ __forceinline__ void ln_row(float (&v)[32], const float* __restrict__ g, const float* __restrict__ bta, int lane) {
  float2 s2 = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < 16; ++j) s2 = __fadd2_rn(s2, make_float2(v[2 * j], v[2 * j + 1]));
  const float mean = warp_sum(s2.x + s2.y) * (1.0f / kDModel);
  const float2 nm = make_float2(-mean, -mean);
  float2 q2 = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 d = __fadd2_rn(make_float2(v[2 * j], v[2 * j + 1]), nm);
    q2 = __ffma2_rn(d, d, q2);
  }
  const float var = warp_sum(q2.x + q2.y) * (1.0f / kDModel);
  const float inv = 1.0f / sqrtf(var + 1e-5f);
  const float2 inv2 = make_float2(inv, inv);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = i * 128 + lane * 4;
    const float4 gg = *reinterpret_cast<const float4*>(g + col);
    const float4 bb = *reinterpret_cast<const float4*>(bta + col);
    const float2 t0 = __fmul2_rn(__fadd2_rn(make_float2(v[4 * i + 0], v[4 * i + 1]), nm), inv2);
    const float2 t1 = __fmul2_rn(__fadd2_rn(make_float2(v[4 * i + 2], v[4 * i + 3]), nm), inv2);
    const float2 r0 = __ffma2_rn(t0, make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
    const float2 r1 = __ffma2_rn(t1, make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
    v[4 * i + 0] = r0.x; v[4 * i + 1] = r0.y; v[4 * i + 2] = r1.x; v[4 * i + 3] = r1.y;
  }
}
#endif

// Relative-position multi-head attention over [ring cache (256) || new rows (Tq)] for every (entry, head).
struct AttnArgs {
  const float* q;          // [M,1024] f32
  const void* kring;       // this layer's K^T ring
  const void* vring;       // this layer's V ring
  const void* ppos_t;      // this layer's projected position table, transposed [head][d][kPosRows]
  int kv_f32;              // element type of the three above (1: f32, 0: bf16)
  const float* bias_u;     // [8,128]
  const float* bias_v;
  ActOut ctx;              // [M,1024]
};
void launch_attention(const BatchDev& b, const AttnArgs& a, cudaStream_t st);

// bf16 tensor-core version (attn_mma.cu): K and V rings both in the natural layout [layer][slot][kRingCap][1024] bf16, fetched by
// TMA through `map_k` / `map_v` (2-D maps over [layers*slots*kRingCap, 1024], box 96 x 64, 128-byte swizzle).
struct AttnMmaArgs {
  const __nv_bfloat16* q_bf16;   // [2][Mcap,1024] bf16: plane 0 = q + pos_bias_u, plane 1 (q_plane elements further) = q + pos_bias_v
  long long q_plane;
  const __nv_bfloat16* g_pos;    // position scores (q + pos_bias_v) . P[r], bf16 [M][8 heads][kPosRowsPad] (batched GEMM, engine.cu)
  ActOut ctx;                    // [M,1024] bf16
  const void* map_k;             // host pointers to 128-byte CUtensorMap objects
  const void* map_v;
  const void* map_k32 = nullptr; // the same rings with 32-row and 8-row boxes: blocks that hold invalid ring slots are fetched as their
  const void* map_v32 = nullptr; // valid 8-key groups only (attn_mma.cu, PARAKEET_B200_ATTN_TRIM); null -> whole 96-key blocks
  const void* map_k8 = nullptr;
  const void* map_v8 = nullptr;
  int layer;
  int n_slots;
  int evict_first = 0;           // set by launch_attention_mma: K / V blocks are loaded with an L2 evict-first policy
  int trim = 0;                  // set by launch_attention_mma: fetch the valid 8-key groups of partly valid blocks only
};
void launch_attention_mma(const BatchDev& b, const AttnMmaArgs& a, cudaStream_t st);

// Conv-module middle: depthwise k=9 over [time cache(4) | c(Tq) | 0000], folded BatchNorm, SiLU; updates the time cache.
struct DwConvArgs {
  const float* c;          // post-GLU activations f32 [M,1024] (precise mode), or
  const __nv_bfloat16* c_bf16;   // ... bf16 [M,1024] (bf16 mode); exactly one of the two is non-null
  float* cache_tm;         // this layer's time cache [slot][1024][4]  (stride between slots passed separately)
  long long slot_stride;   // floats between consecutive slots of cache_tm
  const float* w;          // [1024,9] depthwise weights with BN scale folded in
  const float* bias;       // [1024] folded BN offset
  ActOut out;              // [M,1024]
  const float* wt = nullptr;   // optional transposed copy of w, [9][1024] (dwconv4_kernel: coalesced float4 tap loads)
};
void launch_dwconv(const BatchDev& b, const DwConvArgs& a, cudaStream_t st);

// encoder_output [B,1024,valid] <- first `valid_out` rows of each entry (transposed), zero-filled past qlen
void launch_gather_output(const BatchDev& b, const float* x, float* enc_out /*[B,1024,out_T]*/, int out_T, cudaStream_t st);

// ---- state import/export at the contract layouts (cache carry-over across the C ABI) ----
// acache ring rows (logical 0..255) -> bf16 hi/lo operand rows, for re-projecting K/V after an import
void launch_acache_to_act(const void* acache_layer, int is_f32, const int* slots, const int* heads, int n, ActOut a, cudaStream_t st);
// contract cache_last_channel [n,L,256,1024] f32 (host layout, on device) <-> acache rings
void launch_acache_import(void* acache_layer, int is_f32, const int* slots, const int* heads, int n, const float* src,
                          long long src_entry_stride, cudaStream_t st);
void launch_acache_export(const void* acache_layer, int is_f32, const int* slots, const int* heads, int n, float* dst,
                          long long dst_entry_stride, cudaStream_t st);

}  // namespace pkb
