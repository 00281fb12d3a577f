/*
 * parakeet_trt.h -- the drop-in C ABI of libparakeet_trt.so (B200-native build).
 *
 * Binary-compatible with the reference interface /root/reference/cpp/include/parakeet_trt.h:12-46, which is what
 * rust/parakeet_trt_sys (bindgen, build.rs:8-22) and rust/parakeet_trt/src/lib.rs:24-115 bind.  Same six entry
 * points, same struct layouts, same enum values, same error conventions; behind it sit hand-written sm_100a kernels
 * instead of three TensorRT engines.  Differences a caller can observe:
 *   - <model_dir> must hold `weights.bin` (+ `vocab.txt`) instead of {encoder,predictor,joint}.engine;
 *   - `use_fp16` selects the arithmetic: true  = bf16 tensor-core operands (fast),
 *                                        false = split bf16 hi+lo operands, fp32-grade (precise).
 *     (The reference stores the flag and never consults it: cpp/src/parakeet_trt.cpp:1660, 2314.)
 * Additive entry points (batched engine, audio input, tensor-level calls) live in parakeet_b200.h.
 */
#ifndef PARAKEET_TRT_H
#define PARAKEET_TRT_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Event kinds delivered by parakeet_poll_event (reference header :12-16). */
typedef enum {
    PARAKEET_EVENT_PARTIAL_TEXT = 0,
    PARAKEET_EVENT_FINAL_TEXT = 1,
    PARAKEET_EVENT_ERROR = 2
} ParakeetEventType;

/* `text` / `error_message` are owned by the session, never NULL, and stay valid until the next poll / reset /
 * destroy on that session (reference :18-23; cpp/src/parakeet_trt.cpp:3866-3872).  segment_id is always 0. */
typedef struct {
    ParakeetEventType type;
    int32_t segment_id;
    const char* text;
    const char* error_message;
} ParakeetEvent;

typedef struct ParakeetSession ParakeetSession;

/* `model_dir` is copied; the caller keeps ownership (reference :27-31). */
typedef struct {
    const char* model_dir;
    int32_t device_id;
    bool use_fp16;
} ParakeetConfig;

/* NULL on any failure (NULL config / model_dir, missing files, CUDA error); the reason goes to stderr
 * (reference cpp/src/parakeet_trt.cpp:1700-1843). */
ParakeetSession* parakeet_create_session(const ParakeetConfig* config);

/* NULL-safe (reference :1846). */
void parakeet_destroy_session(ParakeetSession* session);

/* NULL-safe.  Zeroes encoder caches and predictor state, re-primes the predictor with <|startoftranscript|>, <|en|>,
 * clears the accumulated tokens and drains the event queue (reference :1858-1949). */
void parakeet_reset_utterance(ParakeetSession* session);

/* features: bins-major [128, num_frames] contiguous f32, features[m * num_frames + t]; borrowed for the call only.
 * One call == one encoder chunk (pushes above 256 frames are split at 256, reference :1982-2011); all decoding for
 * the pushed frames has completed on return.  Returns 0 ok (also for num_frames == 0), -1 NULL arguments,
 * -2 runtime error (an ERROR event carrying the message is queued) (reference :1967-1969, 3850-3857).
 * Streaming encoder (the default): a chunk must hold 33..256 frames -- 8x subsampling leaves 5 tokens of 33 frames, of which
 * drop_extra_pre_encoded = 2 are dropped and valid_out_len = 3 are decoded; a shorter push returns -2 and processes nothing.
 * All sessions of a process that share (model_dir, device, options) are streams of ONE engine (one copy of the weights); pushes
 * that arrive together from several threads are served by one batched pass (PARAKEET_B200_MAX_SESSIONS slots per engine, default 8;
 * PARAKEET_B200_COALESCE_US = how long a push waits for its siblings, default 200). */
int parakeet_push_features(ParakeetSession* session, const float* features, size_t num_frames);

/* NULL-safe; `id` is copied.  Context shows up in diagnostics only (reference :1951-1965). */
void parakeet_set_debug_context(ParakeetSession* session,
                                const char* id,
                                uint64_t utt_seq,
                                uint64_t audio_chunk_idx,
                                uint64_t feature_idx);

/* FIFO pop; false when the queue is empty or an argument is NULL (reference :3860-3876).
 * May be called from another thread than the pushing one. */
bool parakeet_poll_event(ParakeetSession* session, ParakeetEvent* event);

#ifdef __cplusplus
}
#endif

#endif /* PARAKEET_TRT_H */
