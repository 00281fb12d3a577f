// frontend.h -- host interface of the log-mel frontend kernels (see frontend.cu).
#pragma once
#include "common.cuh"

namespace pkb {

// One contiguous piece of audio whose complete frames are written to `out`.
struct FrontSegment {
  long long audio_off;  // first sample of frame 0 inside the audio buffer (must be even: float2 loads)
  long long out_off;    // float offset of the destination (row 0) inside the output buffer
  int out_stride;       // floats per output row (128 for [T,128])
  int ring_cap;         // 0: rows are linear (row = t); >0: row = (frame0 + t) % ring_cap  (per-stream feature ring)
  int frame0;
  int norm_off;         // -1: none; else float offset into `stats` of [mean[128], std[128]] applied on the fly
};

class Frontend {
 public:
  Frontend();
  ~Frontend();
  Frontend(const Frontend&) = delete;
  // frame_prefix[s] = index of segment s's first frame in the global frame numbering; total_frames = sum.
  void logmel(const float* d_audio, const FrontSegment* d_segs, const int* d_frame_prefix, int n_segs, int total_frames,
              float* d_out, const float* d_stats, int sm_count, cudaStream_t st);
  // utterance-level mean/std over linear [T,128] outputs; stats layout [n_segs][2][128]
  void per_feature_stats(const float* d_feat, const FrontSegment* d_segs, const int* d_frames, int n_segs, float* d_stats,
                         cudaStream_t st);
  void apply_norm(float* d_feat, const FrontSegment* d_segs, const int* d_frames, int n_segs, int max_frames,
                  const float* d_stats, cudaStream_t st);
  // C-ABI features [128,T] bins-major -> ring rows (frame0+t) % ring_cap, frames-major
  void bins_to_frames(const float* d_src, int T, float* d_ring, int ring_cap, int frame0, cudaStream_t st);

 private:
  void* tables_ = nullptr;
};

}  // namespace pkb
