/*
 * parakeet_b200.h -- ADDITIVE C ABI of libparakeet_trt.so (B200-native build).  Nothing here exists in the reference;
 * these entry points are what BASELINE.json's multi-stream / audio-input configurations need on top of the six
 * legacy functions of parakeet_trt.h (SURVEY.md section 8b "Additive API").  Plain C types only.
 *
 *   (i)   a batched engine: one per GPU, owns the weights and a table of stream slots; pkb_engine_step() advances every
 *         stream that has a pending chunk in ONE batched pass (log-mel frontend -> cache-aware FastConformer chunk ->
 *         TDT greedy decode -> cache carry-over).  parakeet_create_session() is a one-stream view of the same engine.
 *   (ii)  audio input: pkb_stream_push_audio() buffers 16 kHz f32 PCM; the step runs the GPU frontend and cuts chunks by
 *         the NeMo cache-aware schedule the reference's harness uses (41 frames, then 57-frame slices shifted by 24:
 *         /root/reference/tools/verify_nemo/streaming_encoder_reference.py:522-550).
 *   (iii) tensor-level calls at the layouts of /root/reference/contracts/parakeet-tdt-0.6b-v3.contract.json
 *         (encoder streaming step :97-159, predictor :169-205, joint :217-241) -- the replacement for running the three
 *         TensorRT engines directly (cpp/src/parakeet_trt.cpp:2443, 2941, 3635); used by the parity tests.
 *
 * Every function returns 0 (or a non-negative count) on success and a negative code on failure; pkb_last_error()
 * returns the message of the calling thread's last failure.  C++ exceptions never cross this boundary.
 */
#ifndef PARAKEET_B200_H
#define PARAKEET_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PkbEngine PkbEngine;

typedef struct {
  const char* model_dir;    /* holds weights.bin + vocab.txt */
  int32_t device_id;        /* one process per GPU (the deployment north_star names): every entry point selects this device for the
                               call, but the kernels' opt-in to > 48 KB of shared memory is configured once per PROCESS, on the
                               device of the first engine -- engines on several devices of one process are not supported */
  int32_t max_streams;      /* stream slots (state is preallocated: ~45 MB per stream in bf16 mode) */
  int32_t precision;        /* 0 = bf16 tensor-core operands; 1 = split bf16 hi+lo operands (fp32-grade) */
  int32_t gemm_backend;     /* 0 = auto (tcgen05 above 16 rows, weight-streaming CUDA-core kernel below), 1 = CUDA cores only,
                               2 = tcgen05 always */
  int32_t contract_cache;   /* 1 = also keep cache_last_channel in contract form (needed by state import/export) */
  int32_t punct_suppression;/* 1 = reference default (leading punctuation-only tokens suppressed, parakeet_trt.cpp:3256) */
  int32_t max_rows;         /* packed encoder rows per batched pass; 0 = 8 * max_streams */
} PkbEngineConfig;

typedef struct {
  int32_t time_idx;         /* encoder frame inside the chunk */
  int32_t token;            /* 8192 = blank */
  int32_t duration;         /* duration-head argmax (0..4) */
} PkbStep;

const char* pkb_last_error(void);
const char* pkb_version(void);

PkbEngine* pkb_engine_create(const PkbEngineConfig* config);       /* NULL on failure */
void pkb_engine_destroy(PkbEngine* engine);
int32_t pkb_engine_num_layers(PkbEngine* engine);
int64_t pkb_engine_kernel_launches(PkbEngine* engine);             /* kernels launched by this engine so far */

/* ---- streams ---- */
int32_t pkb_stream_open(PkbEngine* engine);                          /* -> stream id >= 0 */
int32_t pkb_stream_close(PkbEngine* engine, int32_t stream);
int32_t pkb_stream_reset(PkbEngine* engine, int32_t stream);         /* == parakeet_reset_utterance for that stream */
/* one encoder chunk of T frames, bins-major [128,T] (the parakeet_push_features layout); 33 <= T <= 256 */
int32_t pkb_stream_push_features(PkbEngine* engine, int32_t stream, const float* features, int32_t T);
/* 16 kHz mono f32 PCM; any amount; frames and chunks are formed inside pkb_engine_step() */
int32_t pkb_stream_push_audio(PkbEngine* engine, int32_t stream, const float* pcm, size_t n);
/* per-feature normalisation applied by the GPU frontend: (x - mean[m]) / std[m]; NULLs switch it off */
int32_t pkb_stream_set_feature_norm(PkbEngine* engine, int32_t stream, const float* mean128, const float* std128);
/* Streaming-safe alternative (the reference records its model-matching whole-utterance statistics as "not streaming-safe" and
 * leaves the choice open: /root/reference/docs/DECISION_LOG.md:44-47, 55-58; docs/ARCHITECTURE_RUNTIME.md:48-50): every frame the GPU
 * frontend produces is normalised with the CAUSAL running mean / unbiased std (+1e-5, as rust/features/src/lib.rs:150-158) of the
 * stream's own frames up to and including that frame.  Audio input only; on a freshly opened or reset stream; the mode survives
 * pkb_stream_reset (the statistics restart), pkb_stream_close clears it.  Takes precedence over fixed statistics. */
int32_t pkb_stream_set_feature_norm_running(PkbEngine* engine, int32_t stream, int32_t on);
/* Offline mode = the reference's non-streaming `encoder` engine (contracts/parakeet-tdt-0.6b-v3.contract.json:67-96, selected in
 * cpp/src/parakeet_trt.cpp:1720-1746 when the engine has no cache bindings): every push of 1..256 frames is encoded with full
 * context, no caches are read or carried over, and all of its encoder frames are decoded (predictor state still carries over).
 * Only on a freshly opened or reset stream.  The legacy session enters it with PARAKEET_B200_ENCODER=offline. */
int32_t pkb_stream_set_offline(PkbEngine* engine, int32_t stream, int32_t offline);
/* Batched push: `count` (<= 8192) samples for each of n streams; row i starts at pcm + i*stride floats.  Host memory
 * (pinned memory is copied without a bounce) or, for the _device variant, device memory (e.g. audio already decoded on the
 * GPU).  One H2D copy + one kernel for the whole batch -- the call a multi-stream server makes once per tick. */
int32_t pkb_engine_push_audio_batch(PkbEngine* engine, int32_t n, const int32_t* streams, const float* pcm, int64_t stride,
                                    int32_t count);
int32_t pkb_engine_push_audio_batch_device(PkbEngine* engine, int32_t n, const int32_t* streams, const float* d_pcm,
                                           int64_t stride, int32_t count);
/* advance every stream with a pending chunk by one chunk; returns the number of chunks processed */
int32_t pkb_engine_step(PkbEngine* engine);
int32_t pkb_stream_has_pending(PkbEngine* engine, int32_t stream);

/* ---- measurement hooks ---- */
/* CUDA events recorded on the engine's own stream (device-side timing of a region of steps) */
int32_t pkb_engine_event_record(PkbEngine* engine);                            /* -> event id */
double pkb_engine_event_elapsed_ms(PkbEngine* engine, int32_t ev_a, int32_t ev_b);
/* per-launch CUDA-event timing of the tcgen05 GEMM kernel: enable, run steps, read sum(ms), sum(flops), launches */
int32_t pkb_engine_profile_enable(PkbEngine* engine, int32_t on);
int32_t pkb_engine_profile_read(PkbEngine* engine, double* ms, double* flops, int64_t* launches);
/* same for a kernel class: 0 = tcgen05 GEMM (work = algorithmic FLOPs), 1 = streaming attention, 2 = log-mel frontend, 3 = TDT decode
 * loop (work = algorithmic bytes), 4 = whole-utterance attention (work = algorithmic FLOPs) */
int32_t pkb_engine_profile_read_class(PkbEngine* engine, int32_t cls, double* ms, double* work, int64_t* launches);
/* Steps of a repeated shape run as ONE CUDA graph (encoder chunk + a device-side WHILE node around the decode iteration; set
 * PARAKEET_B200_GRAPH=0 for the launch-by-launch path).  pkb_engine_graphs_built: step shapes currently captured.
 * pkb_engine_decode_loop_stats: device time (CUDA events the graph records around its WHILE node), algorithmic bytes, passes and
 * count of the decode loops run inside graphs since the last reset. */
int32_t pkb_engine_graphs_built(PkbEngine* engine);
/* the reference's PARAKEET_BLANK_PENALTY knob (cpp/src/parakeet_trt.cpp:3175-3178; read from the environment when the engine is
 * created) changed at run time: subtracted from the blank logit before the token argmax of every later decode */
int32_t pkb_engine_set_blank_penalty(PkbEngine* engine, float penalty);
int32_t pkb_engine_decode_loop_stats(PkbEngine* engine, double* ms, double* bytes, int64_t* passes, int64_t* loops, int32_t reset);

/* ---- results ---- */
int32_t pkb_stream_num_tokens(PkbEngine* engine, int32_t stream);
int32_t pkb_stream_tokens(PkbEngine* engine, int32_t stream, int32_t* out, int32_t cap);      /* copies min(n,cap), returns n */
int32_t pkb_stream_last_steps(PkbEngine* engine, int32_t stream, PkbStep* out, int32_t cap);  /* decode trace of the last chunk */
/* Timestamps in the encoder timebase (one encoder frame = 80 ms: 10 ms feature shift x 8 subsampling, docs/ARCHITECTURE_RUNTIME.md:46-47;
 * TDT durations advance encoder frames, docs/DECISION_LOG.md:55-58): the utterance-relative encoder frame of every emitted token
 * (copies min(n,cap), returns n) and the number of encoder frames decoded so far (the live edge). */
int32_t pkb_stream_token_frames(PkbEngine* engine, int32_t stream, int32_t* out, int32_t cap);
int64_t pkb_stream_encoder_frames(PkbEngine* engine, int32_t stream);
/* "Stable prefix + revision window" (MAGNOLIA_INTEGRATION_HANDOFF.md:105-135): how many leading tokens lie at least
 * revision_window_ms behind the live edge, i.e. may be committed as final text; the rest is the revisable suffix. */
int32_t pkb_stream_stable_prefix(PkbEngine* engine, int32_t stream, int32_t revision_window_ms);
int32_t pkb_stream_cache_len(PkbEngine* engine, int32_t stream);                               /* cache_last_channel_len */
int64_t pkb_stream_chunks_done(PkbEngine* engine, int32_t stream);
int32_t pkb_stream_text(PkbEngine* engine, int32_t stream, char* out, int32_t cap);           /* detokenised transcript so far */
int32_t pkb_detokenize(PkbEngine* engine, const int32_t* ids, int32_t n, char* out, int32_t cap);
int32_t pkb_token_is_punct_only(PkbEngine* engine, int32_t id);   /* the predicate of the leading-punctuation suppression */
/* The token table without an engine or a GPU (replaces the reference's Tokenizer, cpp/src/tokenizer.cpp:9-84: constructor = open,
 * decode, is_punct_only, vocab_size).  pkb_vocab_decode copies min(len, cap-1) bytes + NUL and returns the full length. */
typedef struct PkbVocab PkbVocab;
PkbVocab* pkb_vocab_open(const char* vocab_txt_path);              /* NULL on failure (pkb_last_error) */
void pkb_vocab_close(PkbVocab* vocab);
int32_t pkb_vocab_size(const PkbVocab* vocab);
int32_t pkb_vocab_decode(const PkbVocab* vocab, const int32_t* ids, int32_t n, char* out, int32_t cap);
int32_t pkb_vocab_is_punct_only(const PkbVocab* vocab, int32_t id);

/* ---- per-stream state across the ABI (cache carry-over made explicit: checkpoint, migration, functional-mode parity) ----
 * One stream's encoder caches in the contract layout: cache_last_channel [L,256,1024] (valid region = the last
 * cache_len rows), cache_last_time [L,1024,4]; predictor state h,c [2,640], g [640] (+ tokens emitted so far, last token).
 * Needs contract_cache = 1. */
int32_t pkb_stream_import_state(PkbEngine* engine, int32_t stream, const float* cache_last_channel, const float* cache_last_time,
                                int32_t cache_last_channel_len);
int32_t pkb_stream_export_state(PkbEngine* engine, int32_t stream, float* cache_last_channel, float* cache_last_time,
                                int32_t* cache_last_channel_len);
int32_t pkb_stream_set_decoder_state(PkbEngine* engine, int32_t stream, const float* h, const float* c, const float* g,
                                     int32_t n_emitted, int32_t y_id);
int32_t pkb_stream_get_decoder_state(PkbEngine* engine, int32_t stream, float* h, float* c, float* g);

/* ---- tensor-level calls, contract layouts, HOST pointers ---- */
/* encoder_streaming: audio_signal [B,128,T] f32, length [B] i64 (must equal T), cache_last_channel [B,L,256,1024],
 * cache_last_time [B,L,1024,4], cache_last_channel_len [B] i64  ->  encoder_output [B,1024,3], encoded_lengths [B],
 * cache_last_channel_out, cache_last_time_out (same shapes), cache_last_channel_len_out [B]. */
int32_t pkb_encoder_streaming_step(PkbEngine* engine, int32_t B, int32_t T, const float* audio_signal, const int64_t* length,
                                   const float* cache_last_channel, const float* cache_last_time,
                                   const int64_t* cache_last_channel_len, float* encoder_output, int64_t* encoded_lengths,
                                   float* cache_last_channel_out, float* cache_last_time_out,
                                   int64_t* cache_last_channel_len_out);
/* offline encoder: audio_signal [B,128,T] f32 (T <= 256), length [B] i64 (must equal T) -> encoder_output [B,1024,T_enc],
 * encoded_lengths [B]; T_enc = three times floor((L-1)/2)+1. */
int32_t pkb_encoder_offline_step(PkbEngine* engine, int32_t B, int32_t T, const float* audio_signal, const int64_t* length,
                                 float* encoder_output, int64_t* encoded_lengths);
/* Whole-utterance offline path: what the reference's non-streaming `encoder` graph computes when its time axis is dynamic
 * (contracts/parakeet-tdt-0.6b-v3.contract.json:67-96 -- the ONNX export run by tools/onnxruntime, BASELINE configs 1 and 5;
 * the TensorRT build of the same graph is capped at T <= 256 by its profile, contract.json:284-287), followed by the greedy TDT
 * loop of cpp/src/parakeet_trt.cpp:2914-3676 over every encoder frame.  n utterances, utterance i bound to the freshly opened /
 * reset stream streams[i].  Input: either audio (audio[i]: n_samples[i] samples of 16 kHz f32 PCM; log-mel and, if
 * per_feature_norm != 0, the whole-utterance normalisation of rust/features run on the GPU) or features (features[i]: n_frames[i]
 * frames, [128,T] bins-major if bins_major != 0 else [T,128]); the other pair is NULL.  Self-attention spans ALL frames of an
 * utterance; the sum of encoder frames (T/8 each) must fit the engine's max_rows.  encoder_output (NULL, or per-utterance
 * pointers, each NULL or [1024, T_enc_i] f32) receives the encoder output in the contract layout; decode == 1 fills
 * pkb_stream_tokens / pkb_stream_last_steps of each stream before returning; decode == 2 only parks the utterances' encoder
 * rows on the device: pkb_offline_decode_pending() then decodes ALL parked utterances of the engine in one batched TDT loop
 * (clips encoded in several memory-sized groups share one decode pass); it returns the number of utterances decoded.
 * Host pointers. */
int32_t pkb_offline_utterances(PkbEngine* engine, int32_t n, const int32_t* streams, const float* const* audio,
                               const size_t* n_samples, int32_t per_feature_norm, const float* const* features,
                               const int32_t* n_frames, int32_t bins_major, float* const* encoder_output, int32_t decode);
int32_t pkb_offline_decode_pending(PkbEngine* engine);
/* Host restatement hook of the streaming attention's load schedule (no GPU needed): bit q of the result is set when the 8-slot group
 * [8q, 8q+8) of a 288-slot K / V ring holds at least one valid key for an entry whose logical position 0 sits at physical slot `head`,
 * with `len` cached rows and `qlen` new rows -- the groups the kernel fetches.  0 for arguments outside the ring geometry. */
uint64_t pkb_debug_ring_valid_groups(int32_t head, int32_t len, int32_t qlen);
/* encoder frames produced for T feature frames: three times floor((L-1)/2)+1 */
int32_t pkb_encoded_length(int32_t n_frames);
/* predictor: y [B,1] i64, h,c [2,B,640] -> g [B,640,1], h_out,c_out [2,B,640] */
int32_t pkb_predictor_step(PkbEngine* engine, int32_t B, const int64_t* y, const float* h, const float* c, float* g, float* h_out,
                           float* c_out);
/* joint: encoder_output [B,1024,T], predictor_output [B,640,U] -> joint_output [B,T,U,8198] raw logits */
int32_t pkb_joint_step(PkbEngine* engine, int32_t B, int32_t T, int32_t U, const float* encoder_output,
                       const float* predictor_output, float* joint_output);
/* GPU log-mel frontend on host buffers: pcm[n] -> frames-major [n_frames,128]; returns n_frames (or < 0).
 * per_feature_norm != 0 applies the whole-utterance mean/std normalisation of rust/features (lib.rs:127-172). */
int64_t pkb_logmel(PkbEngine* engine, const float* pcm, size_t n, float* out, size_t out_cap_floats, int32_t per_feature_norm);
/* Stand-alone GPU frontend (no model weights): what a caller of the legacy parakeet_push_features() uses in place of
 * rust/features (LogMelExtractor::compute, lib.rs:66-120).  pcm[n] (host, 16 kHz f32) -> frames-major [n_frames,128] log-mel,
 * no normalisation; returns n_frames or < 0. */
typedef struct PkbFrontend PkbFrontend;
PkbFrontend* pkb_frontend_create(int32_t device_id);                 /* NULL on failure */
void pkb_frontend_destroy(PkbFrontend* frontend);
int64_t pkb_frontend_logmel(PkbFrontend* frontend, const float* pcm, size_t n, float* out, size_t out_cap_floats);
/* C[M,N] = A[M,K] (f32) x W[N,K]^T (bf16 bits) through one GEMM backend (0 = CUDA cores, 1 = tcgen05): kernel validation */
int32_t pkb_gemm_test(PkbEngine* engine, int32_t backend, int32_t M, int32_t N, int32_t K, const float* A, const uint16_t* W,
                      float* C);

#ifdef __cplusplus
}
#endif

#endif /* PARAKEET_B200_H */
