/*
 * trt_asr.h -- the reference's secondary ("v2") C ABI, kept as a thin shim over parakeet_trt.h.
 * Binary-compatible with /root/reference/cpp/include/trt_asr.h:42-53 (implementation there: cpp/src/trt_asr.cpp:42-132).
 */
#ifndef TRT_ASR_H
#define TRT_ASR_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TrtAsrSession TrtAsrSession;

typedef struct {
  const char* model_dir;
  int32_t device_id;
  bool use_fp16;
} TrtAsrConfig;

typedef enum {
  TRT_ASR_EVENT_TOKEN = 0,         /* declared by the reference, never produced (trt_asr.cpp:92-132) */
  TRT_ASR_EVENT_PARTIAL_TEXT = 1,
  TRT_ASR_EVENT_FINAL_TEXT = 2,
  TRT_ASR_EVENT_ERROR = 3,
} TrtAsrEventType;

typedef struct {
  TrtAsrEventType type;
  int32_t segment_id;
  int32_t token_id;                /* always -1 */
  const char* text;                /* owned by the session; valid until the next poll / reset / destroy */
  const char* error_message;
} TrtAsrEvent;

TrtAsrSession* trt_asr_create_session(const TrtAsrConfig* config);
void trt_asr_destroy_session(TrtAsrSession* session);
void trt_asr_reset_session(TrtAsrSession* session);

/* features: [128, T] bins-major; `length` = valid frames (<= T).  Returns -1 on bad arguments, else the
 * parakeet_push_features code.  The f16 variant widens 128*T IEEE halfs on the host (trt_asr.cpp:82-90). */
int trt_asr_push_features_f16(TrtAsrSession* session, const uint16_t* features_f16, int32_t T, int32_t length);
int trt_asr_push_features_f32(TrtAsrSession* session, const float* features_f32, int32_t T, int32_t length);

bool trt_asr_poll_event(TrtAsrSession* session, TrtAsrEvent* out_event);

#ifdef __cplusplus
}
#endif

#endif /* TRT_ASR_H */
