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
  int run_state;        // 0: none; else 1 + float offset into the running-statistics state [count, mean[128], M2[128]] (running_norm)
};

// One staged feature push: T frames, bins-major [128,T] at float offset src_off of the staging area, destined for ring rows
// (frame0 + t) % ring_cap of the ring starting at float offset ring_off.
struct FeatPush {
  long long src_off;
  long long ring_off;
  int T;
  int frame0;
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
  // Streaming-safe per-feature normalisation (docs/DECISION_LOG.md:44-47 flags the model's whole-utterance statistics as "not
  // streaming-safe"): every new frame is normalised with the CAUSAL running mean / unbiased std of its own stream up to and including
  // that frame (Welford update in frame order, std + 1e-5 like lib.rs:150-158; the first frame of a stream, which has no variance yet,
  // becomes 0).  Runs over the frames one logmel() launch wrote: segment s holds frame_prefix[s+1] - frame_prefix[s] new frames.
  void running_norm(float* d_out, const FrontSegment* d_segs, const int* d_frame_prefix, int n_segs, float* d_state, cudaStream_t st);
  // C-ABI features [128,T] bins-major -> ring rows (frame0+t) % ring_cap, frames-major
  void bins_to_frames(const float* d_src, int T, float* d_ring, int ring_cap, int frame0, cudaStream_t st);
  // the same for n staged pushes in one launch (grid.z = push)
  void bins_to_frames_batch(const float* d_stage, const FeatPush* d_push, int n, int max_T, float* d_rings, int ring_cap, cudaStream_t st);

 private:
  void* tables_ = nullptr;
};

}  // namespace pkb
