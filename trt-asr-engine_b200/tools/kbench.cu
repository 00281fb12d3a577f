// kbench.cu -- kernel micro-benchmarks on the GPU box (CUDA events, warm, back-to-back launches).  Development tool:
// not part of libparakeet_trt.so.    usage: kbench gemm M N K [iters] [epi] [interleave_ln]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../csrc/enc_kernels.cuh"
#include "../csrc/gemm.h"

using namespace pkb;

static float* dalloc_f(size_t n) { float* p; PKB_CUDA(cudaMalloc(&p, n * 4)); PKB_CUDA(cudaMemset(p, 0, n * 4)); return p; }

// `kbench clusters`: how many thread-block clusters of 2 / 4 / 8 CTAs with a GEMM-sized shared-memory footprint (one CTA per SM) the
// device can hold at once -- decides whether operand multicast over 4-CTA clusters can keep all 148 SMs busy.
__global__ void probe_kernel(int* p) { if (p) *p = 0; }
static int probe_clusters() {
  PKB_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PKB_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs)%s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (e != cudaSuccess) cudaGetLastError();
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && strcmp(argv[1], "clusters") == 0) return probe_clusters();
  if (argc < 5 || strcmp(argv[1], "gemm") != 0) { printf("usage: kbench gemm M N K [iters] [epi: f32|resadd|silu|glu] [ln 0/1]\n"); return 1; }
  const int M = atoi(argv[2]), N = atoi(argv[3]), K = atoi(argv[4]);
  const int iters = argc > 5 ? atoi(argv[5]) : 50;
  const char* epi = argc > 6 ? argv[6] : "f32";
  const int with_ln = argc > 7 ? atoi(argv[7]) : 0;
  const int rot = argc > 8 ? atoi(argv[8]) : 1;      // number of distinct weight copies rotated through (rot * N * K * 2 B > L2: cold weights)
  cudaStream_t st;
  PKB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const size_t rows_a = ((size_t)M + 127) / 128 * 128, rows_w = ((size_t)N + 127) / 128 * 128;
  __nv_bfloat16 *A, *W, *act;
  PKB_CUDA(cudaMalloc(&A, rows_a * K * 2));
  PKB_CUDA(cudaMalloc(&W, (size_t)rot * rows_w * K * 2));
  PKB_CUDA(cudaMalloc(&act, rows_a * (size_t)N * 2));
  std::vector<uint16_t> h(rows_a * K);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 0x3c00 + (uint16_t)(i * 2654435761u >> 24);   // small positive bf16 values
  PKB_CUDA(cudaMemcpy(A, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  h.assign(rows_w * K, 0);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 0x3800 + (uint16_t)(i * 40503u >> 8 & 0xff);
  for (int r = 0; r < rot; ++r) PKB_CUDA(cudaMemcpy(W + (size_t)r * rows_w * K, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  float* out = dalloc_f((size_t)M * N);
  float* x = dalloc_f((size_t)M * 1024);
  float* g1 = dalloc_f(1024);
  __nv_bfloat16* lnA;
  PKB_CUDA(cudaMalloc(&lnA, rows_a * 1024 * 2));
  TensorMap ma;
  std::vector<TensorMap> mws(rot), mws32(rot);
  make_tensor_map_2d(&ma, A, rows_a, K, K, 128);
  for (int r = 0; r < rot; ++r) {
    make_tensor_map_2d(&mws[r], W + (size_t)r * rows_w * K, rows_w, K, K, 128);
    make_tensor_map_2d(&mws32[r], W + (size_t)r * rows_w * K, rows_w, K, K, 32);
  }
  GemmArgs g;
  g.A = A; g.lda = K; g.W = W; g.M = M; g.N = N; g.K = K;
  if (!strcmp(epi, "resadd")) { g.epi.mode = EPI_RESADD_F32; g.epi.out_f32 = out; g.epi.ldo = N; g.epi.scale = 0.5f; }
  else if (!strcmp(epi, "silu")) { g.epi.mode = EPI_SILU_ACT; g.epi.out_act = act; g.epi.lda_out = N; }
  else if (!strcmp(epi, "glu")) { g.epi.mode = EPI_GLU_F32; g.epi.out_f32 = out; g.epi.ldo = N / 2; }
  else if (!strncmp(epi, "partial", 7)) {      // partial<splits>[p][b]: split-K partial sums; p = split chosen for the CTA-pair kernel; b = bf16 partials
    const int sp = epi[7] ? epi[7] - '0' : 1;
    float* ws = dalloc_f((size_t)sp * M * N);
    g.epi.mode = EPI_PARTIAL_F32; g.epi.out_f32 = ws; g.epi.ldo = N; g.epi.splits = sp; g.epi.part_rows = M;
    g.epi.pair_split = strchr(epi + 7, 'p') ? 1 : 0;
    g.epi.part_bf16 = strchr(epi + 7, 'b') ? 1 : 0;
  }
  else { g.epi.mode = EPI_F32; g.epi.out_f32 = out; g.epi.ldo = N; }
  PKB_CUDA(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  PKB_CUDA(cudaEventCreate(&e0));
  PKB_CUDA(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {
    for (int i = 0; i < 5; ++i) gemm_tc(g, ma, mws[0], st);
    PKB_CUDA(cudaStreamSynchronize(st));
    PKB_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) {
      g.W = W + (size_t)(i % rot) * rows_w * K;
      g.map_w32 = &mws32[i % rot];
      gemm_tc(g, ma, mws[i % rot], st);
      if (with_ln) launch_layernorm(x, M, g1, g1, nullptr, nullptr, 0, ActOut{lnA, 1024, 0}, nullptr, st);
    }
    PKB_CUDA(cudaEventRecord(e1, st));
    PKB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PKB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double us = 1e3 * ms / iters;
    printf("gemm M=%d N=%d K=%d epi=%s ln=%d rot=%d: %.2f us/iter  %.1f TFLOP/s\n", M, N, K, epi, with_ln, rot, us, 2.0 * M * N * K / us * 1e-6);
  }
  return 0;
}
