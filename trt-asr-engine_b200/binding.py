"""ctypes binding of libparakeet_trt.so -- the host-side mirror of the reference's FFI layers.

`ParakeetSessionSafe` mirrors /root/reference/rust/parakeet_trt/src/lib.rs:24-115 (the safe wrapper over the bindgen'd
/root/reference/rust/parakeet_trt_sys): same method names, argument meaning and error behaviour, so tests written
against it read like the reference's own callers (rust/cli/src/main.rs:289-293, 408-474).
`Engine` exposes the additive batched / audio / tensor-level entry points of include/parakeet_b200.h.

The product path FAILS LOUDLY when the CUDA library is missing or no B200 is present: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libparakeet_trt.so")

PARAKEET_EVENT_PARTIAL_TEXT, PARAKEET_EVENT_FINAL_TEXT, PARAKEET_EVENT_ERROR = 0, 1, 2

LEGACY_SYMBOLS = ["parakeet_create_session", "parakeet_destroy_session", "parakeet_reset_utterance", "parakeet_push_features",
                  "parakeet_set_debug_context", "parakeet_poll_event"]
TRT_ASR_SYMBOLS = ["trt_asr_create_session", "trt_asr_destroy_session", "trt_asr_reset_session", "trt_asr_push_features_f16",
                   "trt_asr_push_features_f32", "trt_asr_poll_event"]
B200_SYMBOLS = ["pkb_last_error", "pkb_version", "pkb_engine_create", "pkb_engine_destroy", "pkb_engine_num_layers",
                "pkb_engine_kernel_launches", "pkb_stream_open", "pkb_stream_close", "pkb_stream_reset", "pkb_stream_push_features",
                "pkb_stream_push_audio", "pkb_stream_set_feature_norm", "pkb_stream_set_feature_norm_running", "pkb_stream_set_offline", "pkb_encoder_offline_step", "pkb_offline_utterances", "pkb_offline_decode_pending", "pkb_encoded_length", "pkb_debug_ring_valid_groups", "pkb_engine_push_audio_batch",
                "pkb_engine_push_audio_batch_device", "pkb_engine_event_record", "pkb_engine_event_elapsed_ms",
                "pkb_engine_profile_enable", "pkb_engine_profile_read", "pkb_engine_profile_read_class", "pkb_engine_graphs_built", "pkb_engine_set_blank_penalty",
                "pkb_engine_decode_loop_stats", "pkb_engine_step", "pkb_stream_has_pending",
                "pkb_stream_num_tokens", "pkb_stream_tokens", "pkb_stream_token_frames", "pkb_stream_encoder_frames", "pkb_stream_stable_prefix", "pkb_stream_last_steps", "pkb_stream_cache_len",
                "pkb_stream_chunks_done", "pkb_stream_text", "pkb_detokenize", "pkb_token_is_punct_only", "pkb_vocab_open", "pkb_vocab_close",
                "pkb_vocab_size", "pkb_vocab_decode", "pkb_vocab_is_punct_only", "pkb_stream_import_state", "pkb_stream_export_state",
                "pkb_stream_set_decoder_state", "pkb_stream_get_decoder_state", "pkb_encoder_streaming_step", "pkb_predictor_step",
                "pkb_joint_step", "pkb_logmel", "pkb_gemm_test", "pkb_frontend_create", "pkb_frontend_destroy", "pkb_frontend_logmel"]


class ParakeetConfig(C.Structure):
    _fields_ = [("model_dir", C.c_char_p), ("device_id", C.c_int32), ("use_fp16", C.c_bool)]


class ParakeetEvent(C.Structure):
    _fields_ = [("type", C.c_int), ("segment_id", C.c_int32), ("text", C.c_char_p), ("error_message", C.c_char_p)]


class TrtAsrEvent(C.Structure):
    _fields_ = [("type", C.c_int), ("segment_id", C.c_int32), ("token_id", C.c_int32), ("text", C.c_char_p),
                ("error_message", C.c_char_p)]


class PkbEngineConfig(C.Structure):
    _fields_ = [("model_dir", C.c_char_p), ("device_id", C.c_int32), ("max_streams", C.c_int32), ("precision", C.c_int32),
                ("gemm_backend", C.c_int32), ("contract_cache", C.c_int32), ("punct_suppression", C.c_int32),
                ("max_rows", C.c_int32)]


class PkbStep(C.Structure):
    _fields_ = [("time_idx", C.c_int32), ("token", C.c_int32), ("duration", C.c_int32)]


_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen the library and declare every prototype.  Raises if the .so is absent (run __graft_entry__.build())."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    lib = C.CDLL(p)
    vp, fp, ip, lp = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    lib.parakeet_create_session.restype = vp
    lib.parakeet_create_session.argtypes = [C.POINTER(ParakeetConfig)]
    lib.parakeet_destroy_session.argtypes = [vp]
    lib.parakeet_destroy_session.restype = None
    lib.parakeet_reset_utterance.argtypes = [vp]
    lib.parakeet_reset_utterance.restype = None
    lib.parakeet_push_features.argtypes = [vp, fp, C.c_size_t]
    lib.parakeet_push_features.restype = C.c_int
    lib.parakeet_set_debug_context.argtypes = [vp, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.parakeet_set_debug_context.restype = None
    lib.parakeet_poll_event.argtypes = [vp, C.POINTER(ParakeetEvent)]
    lib.parakeet_poll_event.restype = C.c_bool
    lib.trt_asr_create_session.restype = vp
    lib.trt_asr_create_session.argtypes = [C.POINTER(ParakeetConfig)]
    lib.trt_asr_destroy_session.argtypes = [vp]
    lib.trt_asr_destroy_session.restype = None
    lib.trt_asr_reset_session.argtypes = [vp]
    lib.trt_asr_reset_session.restype = None
    lib.trt_asr_push_features_f32.argtypes = [vp, fp, C.c_int32, C.c_int32]
    lib.trt_asr_push_features_f16.argtypes = [vp, C.POINTER(C.c_uint16), C.c_int32, C.c_int32]
    lib.trt_asr_poll_event.argtypes = [vp, C.POINTER(TrtAsrEvent)]
    lib.trt_asr_poll_event.restype = C.c_bool
    lib.pkb_last_error.restype = C.c_char_p
    lib.pkb_version.restype = C.c_char_p
    lib.pkb_engine_create.restype = vp
    lib.pkb_engine_create.argtypes = [C.POINTER(PkbEngineConfig)]
    lib.pkb_engine_destroy.argtypes = [vp]
    lib.pkb_engine_destroy.restype = None
    lib.pkb_engine_num_layers.argtypes = [vp]
    lib.pkb_engine_kernel_launches.argtypes = [vp]
    lib.pkb_engine_kernel_launches.restype = C.c_int64
    for name in ("pkb_stream_open", "pkb_engine_step"):
        getattr(lib, name).argtypes = [vp]
    for name in ("pkb_stream_close", "pkb_stream_reset", "pkb_stream_has_pending", "pkb_stream_num_tokens", "pkb_stream_cache_len"):
        getattr(lib, name).argtypes = [vp, C.c_int32]
    lib.pkb_stream_chunks_done.argtypes = [vp, C.c_int32]
    lib.pkb_stream_chunks_done.restype = C.c_int64
    lib.pkb_stream_push_features.argtypes = [vp, C.c_int32, fp, C.c_int32]
    lib.pkb_stream_push_audio.argtypes = [vp, C.c_int32, fp, C.c_size_t]
    lib.pkb_stream_set_feature_norm.argtypes = [vp, C.c_int32, fp, fp]
    lib.pkb_engine_push_audio_batch.argtypes = [vp, C.c_int32, ip, C.c_void_p, C.c_int64, C.c_int32]
    lib.pkb_engine_push_audio_batch_device.argtypes = [vp, C.c_int32, ip, C.c_void_p, C.c_int64, C.c_int32]
    lib.pkb_engine_event_record.argtypes = [vp]
    lib.pkb_engine_event_elapsed_ms.argtypes = [vp, C.c_int32, C.c_int32]
    lib.pkb_engine_event_elapsed_ms.restype = C.c_double
    lib.pkb_engine_profile_enable.argtypes = [vp, C.c_int32]
    lib.pkb_engine_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), lp]
    lib.pkb_engine_profile_read_class.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), lp]
    lib.pkb_engine_graphs_built.argtypes = [vp]
    lib.pkb_engine_set_blank_penalty.argtypes = [vp, C.c_float]
    lib.pkb_engine_decode_loop_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), lp, lp, C.c_int32]
    lib.pkb_stream_tokens.argtypes = [vp, C.c_int32, ip, C.c_int32]
    lib.pkb_stream_last_steps.argtypes = [vp, C.c_int32, C.POINTER(PkbStep), C.c_int32]
    lib.pkb_stream_token_frames.argtypes = [vp, C.c_int32, ip, C.c_int32]
    lib.pkb_stream_encoder_frames.argtypes = [vp, C.c_int32]
    lib.pkb_stream_encoder_frames.restype = C.c_int64
    lib.pkb_stream_stable_prefix.argtypes = [vp, C.c_int32, C.c_int32]
    lib.pkb_stream_text.argtypes = [vp, C.c_int32, C.c_char_p, C.c_int32]
    lib.pkb_detokenize.argtypes = [vp, ip, C.c_int32, C.c_char_p, C.c_int32]
    lib.pkb_token_is_punct_only.argtypes = [vp, C.c_int32]
    lib.pkb_vocab_open.restype = vp
    lib.pkb_vocab_open.argtypes = [C.c_char_p]
    lib.pkb_vocab_close.argtypes = [vp]
    lib.pkb_vocab_close.restype = None
    lib.pkb_vocab_size.argtypes = [vp]
    lib.pkb_vocab_decode.argtypes = [vp, ip, C.c_int32, C.c_char_p, C.c_int32]
    lib.pkb_vocab_is_punct_only.argtypes = [vp, C.c_int32]
    lib.pkb_stream_import_state.argtypes = [vp, C.c_int32, fp, fp, C.c_int32]
    lib.pkb_stream_export_state.argtypes = [vp, C.c_int32, fp, fp, ip]
    lib.pkb_stream_set_decoder_state.argtypes = [vp, C.c_int32, fp, fp, fp, C.c_int32, C.c_int32]
    lib.pkb_stream_get_decoder_state.argtypes = [vp, C.c_int32, fp, fp, fp]
    lib.pkb_encoder_streaming_step.argtypes = [vp, C.c_int32, C.c_int32, fp, lp, fp, fp, lp, fp, lp, fp, fp, lp]
    lib.pkb_stream_set_offline.argtypes = [vp, C.c_int32, C.c_int32]
    lib.pkb_stream_set_feature_norm_running.argtypes = [vp, C.c_int32, C.c_int32]
    lib.pkb_encoder_offline_step.argtypes = [vp, C.c_int32, C.c_int32, fp, lp, fp, lp]
    lib.pkb_offline_utterances.argtypes = [vp, C.c_int32, ip, C.POINTER(fp), C.POINTER(C.c_size_t), C.c_int32, C.POINTER(fp), ip, C.c_int32,
                                           C.POINTER(fp), C.c_int32]
    lib.pkb_encoded_length.argtypes = [C.c_int32]
    lib.pkb_debug_ring_valid_groups.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.pkb_debug_ring_valid_groups.restype = C.c_uint64
    lib.pkb_offline_decode_pending.argtypes = [vp]
    lib.pkb_predictor_step.argtypes = [vp, C.c_int32, lp, fp, fp, fp, fp, fp]
    lib.pkb_joint_step.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, fp, fp, fp]
    lib.pkb_logmel.argtypes = [vp, fp, C.c_size_t, fp, C.c_size_t, C.c_int32]
    lib.pkb_logmel.restype = C.c_int64
    lib.pkb_gemm_test.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, fp, C.POINTER(C.c_uint16), fp]
    if path is None:
        _lib = lib
    return lib


def _fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _lptr(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class TranscriptionEvent:
    """rust/parakeet_trt/src/lib.rs:9-22 (enum TranscriptionEvent)."""

    def __init__(self, kind: str, text: str = "", segment_id: int = 0, message: str = ""):
        self.kind, self.text, self.segment_id, self.message = kind, text, segment_id, message

    def __repr__(self):
        return f"TranscriptionEvent({self.kind!r}, text={self.text!r}, message={self.message!r})"


class ParakeetSessionSafe:
    """Mirror of `ParakeetSessionSafe` (rust/parakeet_trt/src/lib.rs:24-115)."""

    def __init__(self, model_dir: str, device_id: int = 0, use_fp16: bool = True):
        self._lib = load_library()
        cfg = ParakeetConfig(model_dir.encode(), device_id, use_fp16)
        self._s = self._lib.parakeet_create_session(C.byref(cfg))
        if not self._s:  # lib.rs:38-40
            raise RuntimeError("Failed to create Parakeet session")

    def reset(self) -> None:
        self._lib.parakeet_reset_utterance(self._s)

    def push_features(self, features_bct_f32: np.ndarray, num_frames: int) -> None:
        """features: bins-major [128, num_frames] f32 (lib.rs:50-66); raises on a non-zero return code."""
        f = np.ascontiguousarray(features_bct_f32, dtype=np.float32)
        rc = self._lib.parakeet_push_features(self._s, _fptr(f), num_frames)
        if rc != 0:
            raise RuntimeError(f"Failed to push features: error code {rc}")

    def set_debug_context(self, id: str, utt_seq: int, audio_chunk_idx: int, feature_idx: int) -> None:
        self._lib.parakeet_set_debug_context(self._s, id.encode(), utt_seq, audio_chunk_idx, feature_idx)

    def poll_event(self) -> Optional[TranscriptionEvent]:
        ev = ParakeetEvent()
        if not self._lib.parakeet_poll_event(self._s, C.byref(ev)):
            return None
        if ev.type == PARAKEET_EVENT_PARTIAL_TEXT:
            return TranscriptionEvent("partial", (ev.text or b"").decode("utf-8", "replace"), ev.segment_id)
        if ev.type == PARAKEET_EVENT_FINAL_TEXT:
            return TranscriptionEvent("final", (ev.text or b"").decode("utf-8", "replace"), ev.segment_id)
        return TranscriptionEvent("error", message=(ev.error_message or b"").decode("utf-8", "replace"))

    def close(self) -> None:
        if self._s:
            self._lib.parakeet_destroy_session(self._s)
            self._s = None

    def __del__(self):  # lib.rs:108-112 (Drop)
        try:
            self.close()
        except Exception:
            pass


class Vocab:
    """The product's token table on its own (pkb_vocab_*; no GPU): mirror of the reference's `Tokenizer` (cpp/include/tokenizer.h)."""

    def __init__(self, vocab_path: str):
        self._lib = load_library()
        self._v = self._lib.pkb_vocab_open(vocab_path.encode())
        if not self._v:
            raise RuntimeError("pkb_vocab_open failed: " + self._lib.pkb_last_error().decode())

    def vocab_size(self) -> int:
        return int(self._lib.pkb_vocab_size(self._v))

    def decode(self, ids) -> str:
        a = np.ascontiguousarray(ids, np.int32)
        n = self._lib.pkb_vocab_decode(self._v, a.ctypes.data_as(C.POINTER(C.c_int32)), a.size, None, 0)
        buf = C.create_string_buffer(n + 1)
        self._lib.pkb_vocab_decode(self._v, a.ctypes.data_as(C.POINTER(C.c_int32)), a.size, buf, n + 1)
        return buf.value.decode("utf-8", "replace")

    def is_punct_only(self, i: int) -> bool:
        return bool(self._lib.pkb_vocab_is_punct_only(self._v, int(i)))

    def close(self):
        if self._v:
            self._lib.pkb_vocab_close(self._v)
            self._v = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """Batched multi-stream engine (include/parakeet_b200.h)."""

    def __init__(self, model_dir: str, device_id: int = 0, max_streams: int = 1, precision: int = 0, gemm_backend: int = 0,
                 contract_cache: int = 1, punct_suppression: int = 1, max_rows: int = 0):
        self._lib = load_library()
        cfg = PkbEngineConfig(model_dir.encode(), device_id, max_streams, precision, gemm_backend, contract_cache,
                              punct_suppression, max_rows)
        self._e = self._lib.pkb_engine_create(C.byref(cfg))
        if not self._e:
            raise RuntimeError("pkb_engine_create failed: " + self._lib.pkb_last_error().decode())
        self.n_layers = self._lib.pkb_engine_num_layers(self._e)

    def _chk(self, rc: int) -> int:
        if rc < 0:
            raise RuntimeError(f"libparakeet_trt error {rc}: " + self._lib.pkb_last_error().decode())
        return rc

    def close(self):
        if self._e:
            self._lib.pkb_engine_destroy(self._e)
            self._e = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # streams
    def open(self) -> int:
        return self._chk(self._lib.pkb_stream_open(self._e))

    def close_stream(self, s: int):
        self._chk(self._lib.pkb_stream_close(self._e, s))

    def reset(self, s: int):
        self._chk(self._lib.pkb_stream_reset(self._e, s))

    def push_features(self, s: int, feats_bins_major: np.ndarray):
        f = np.ascontiguousarray(feats_bins_major, dtype=np.float32)
        assert f.shape[0] == 128
        self._chk(self._lib.pkb_stream_push_features(self._e, s, _fptr(f), f.shape[1]))

    def push_audio(self, s: int, pcm: np.ndarray):
        a = np.ascontiguousarray(pcm, dtype=np.float32)
        self._chk(self._lib.pkb_stream_push_audio(self._e, s, _fptr(a), a.size))

    def set_feature_norm(self, s: int, mean: Optional[np.ndarray], std: Optional[np.ndarray]):
        if mean is None:
            self._chk(self._lib.pkb_stream_set_feature_norm(self._e, s, None, None))
        else:
            m, d = np.ascontiguousarray(mean, np.float32), np.ascontiguousarray(std, np.float32)
            self._chk(self._lib.pkb_stream_set_feature_norm(self._e, s, _fptr(m), _fptr(d)))

    def set_feature_norm_running(self, s: int, on: bool = True):
        """streaming-safe normalisation: causal running mean / std per feature (GPU frontend, audio input)."""
        self._chk(self._lib.pkb_stream_set_feature_norm_running(self._e, s, int(on)))

    def push_audio_batch(self, sids: np.ndarray, host_ptr: int, stride: int, count: int):
        """sids int32 array; host_ptr = address of row 0 (e.g. tensor.data_ptr() of pinned memory)."""
        self._chk(self._lib.pkb_engine_push_audio_batch(self._e, sids.size, sids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                        C.c_void_p(host_ptr), stride, count))

    def push_audio_batch_device(self, sids: np.ndarray, dev_ptr: int, stride: int, count: int):
        self._chk(self._lib.pkb_engine_push_audio_batch_device(self._e, sids.size, sids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                               C.c_void_p(dev_ptr), stride, count))

    def event_record(self) -> int:
        return self._chk(self._lib.pkb_engine_event_record(self._e))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        return float(self._lib.pkb_engine_event_elapsed_ms(self._e, a, b))

    def profile_enable(self, on: bool):
        self._chk(self._lib.pkb_engine_profile_enable(self._e, int(on)))

    def profile_read(self):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        self._chk(self._lib.pkb_engine_profile_read(self._e, C.byref(ms), C.byref(fl), C.byref(n)))
        return ms.value, fl.value, n.value

    def profile_read_class(self, cls: int):
        """cls 0: tcgen05 GEMM (work = FLOPs); 1: attention; 2: log-mel frontend (work = algorithmic bytes)."""
        ms, wk, n = C.c_double(), C.c_double(), C.c_int64()
        self._chk(self._lib.pkb_engine_profile_read_class(self._e, cls, C.byref(ms), C.byref(wk), C.byref(n)))
        return ms.value, wk.value, n.value

    def set_blank_penalty(self, penalty: float):
        self._chk(self._lib.pkb_engine_set_blank_penalty(self._e, float(penalty)))

    def graphs_built(self) -> int:
        return self._chk(self._lib.pkb_engine_graphs_built(self._e))

    def decode_loop_stats(self, reset: bool = True):
        """(ms, algorithmic bytes, passes, loops) of the decode loops run inside step graphs since the last reset."""
        ms, by, p, l = C.c_double(), C.c_double(), C.c_int64(), C.c_int64()
        self._chk(self._lib.pkb_engine_decode_loop_stats(self._e, C.byref(ms), C.byref(by), C.byref(p), C.byref(l), int(reset)))
        return ms.value, by.value, p.value, l.value

    def step(self) -> int:
        return self._chk(self._lib.pkb_engine_step(self._e))

    def has_pending(self, s: int) -> bool:
        return bool(self._chk(self._lib.pkb_stream_has_pending(self._e, s)))

    def tokens(self, s: int) -> List[int]:
        n = self._chk(self._lib.pkb_stream_num_tokens(self._e, s))
        out = np.zeros(max(n, 1), np.int32)
        self._chk(self._lib.pkb_stream_tokens(self._e, s, out.ctypes.data_as(C.POINTER(C.c_int32)), n))
        return out[:n].tolist()

    def last_steps(self, s: int) -> List[Tuple[int, int, int]]:
        cap = 320
        buf = (PkbStep * cap)()
        n = self._chk(self._lib.pkb_stream_last_steps(self._e, s, buf, cap))
        if n > cap:      # a whole-utterance decode trace
            cap = n
            buf = (PkbStep * cap)()
            n = self._chk(self._lib.pkb_stream_last_steps(self._e, s, buf, cap))
        return [(buf[i].time_idx, buf[i].token, buf[i].duration) for i in range(min(n, cap))]

    def token_frames(self, s: int) -> List[int]:
        n = self._chk(self._lib.pkb_stream_token_frames(self._e, s, None, 0))
        out = np.zeros(max(n, 1), np.int32)
        self._chk(self._lib.pkb_stream_token_frames(self._e, s, out.ctypes.data_as(C.POINTER(C.c_int32)), n))
        return out[:n].tolist()

    def encoder_frames(self, s: int) -> int:
        return self._chk(int(self._lib.pkb_stream_encoder_frames(self._e, s)))

    def stable_prefix(self, s: int, revision_window_ms: int) -> int:
        return self._chk(self._lib.pkb_stream_stable_prefix(self._e, s, revision_window_ms))

    def cache_len(self, s: int) -> int:
        return self._chk(self._lib.pkb_stream_cache_len(self._e, s))

    def chunks_done(self, s: int) -> int:
        return int(self._lib.pkb_stream_chunks_done(self._e, s))

    def text(self, s: int) -> str:
        buf = C.create_string_buffer(1 << 16)
        self._chk(self._lib.pkb_stream_text(self._e, s, buf, len(buf)))
        return buf.value.decode("utf-8", "replace")

    def detokenize(self, ids: List[int]) -> str:
        a = np.asarray(ids, np.int32)
        buf = C.create_string_buffer(1 << 16)
        self._chk(self._lib.pkb_detokenize(self._e, a.ctypes.data_as(C.POINTER(C.c_int32)), a.size, buf, len(buf)))
        return buf.value.decode("utf-8", "replace")

    def token_is_punct_only(self, i: int) -> bool:
        return bool(self._chk(self._lib.pkb_token_is_punct_only(self._e, int(i))))

    def kernel_launches(self) -> int:
        return int(self._lib.pkb_engine_kernel_launches(self._e))

    # per-stream state across the ABI
    def import_state(self, s: int, cache_ch: np.ndarray, cache_tm: np.ndarray, cache_len: int):
        cc, ct = np.ascontiguousarray(cache_ch, np.float32), np.ascontiguousarray(cache_tm, np.float32)
        assert cc.shape == (self.n_layers, 256, 1024) and ct.shape == (self.n_layers, 1024, 4)
        self._chk(self._lib.pkb_stream_import_state(self._e, s, _fptr(cc), _fptr(ct), int(cache_len)))

    def export_state(self, s: int):
        cc, ct = np.zeros((self.n_layers, 256, 1024), np.float32), np.zeros((self.n_layers, 1024, 4), np.float32)
        ln = C.c_int32()
        self._chk(self._lib.pkb_stream_export_state(self._e, s, _fptr(cc), _fptr(ct), C.byref(ln)))
        return cc, ct, ln.value

    def set_decoder_state(self, s: int, h: np.ndarray, c: np.ndarray, g: np.ndarray, n_emitted: int, y_id: int):
        h, c, g = (np.ascontiguousarray(a, np.float32) for a in (h, c, g))
        assert h.size == 1280 and c.size == 1280 and g.size == 640
        self._chk(self._lib.pkb_stream_set_decoder_state(self._e, s, _fptr(h), _fptr(c), _fptr(g), n_emitted, y_id))

    def get_decoder_state(self, s: int):
        h, c, g = np.zeros((2, 640), np.float32), np.zeros((2, 640), np.float32), np.zeros(640, np.float32)
        self._chk(self._lib.pkb_stream_get_decoder_state(self._e, s, _fptr(h), _fptr(c), _fptr(g)))
        return h, c, g

    # tensor-level calls (contract layouts)
    def encoder_streaming_step(self, audio_signal, length, cache_last_channel, cache_last_time, cache_last_channel_len):
        a = np.ascontiguousarray(audio_signal, np.float32)
        B, _, T = a.shape
        ln = np.ascontiguousarray(length, np.int64)
        cc = np.ascontiguousarray(cache_last_channel, np.float32)
        ct = np.ascontiguousarray(cache_last_time, np.float32)
        cl = np.ascontiguousarray(cache_last_channel_len, np.int64)
        enc = np.zeros((B, 1024, 3), np.float32)
        el = np.zeros(B, np.int64)
        cco, cto, clo = np.zeros_like(cc), np.zeros_like(ct), np.zeros(B, np.int64)
        self._chk(self._lib.pkb_encoder_streaming_step(self._e, B, T, _fptr(a), _lptr(ln), _fptr(cc), _fptr(ct), _lptr(cl), _fptr(enc),
                                                       _lptr(el), _fptr(cco), _fptr(cto), _lptr(clo)))
        return enc, el, cco, cto, clo

    def set_offline(self, s: int, offline: bool = True):
        self._chk(self._lib.pkb_stream_set_offline(self._e, s, int(offline)))

    def encoder_offline_step(self, audio_signal, length):
        """audio_signal [B,128,T] (T <= 256), length [B] -> (encoder_output [B,1024,T_enc], encoded_lengths [B])."""
        a = np.ascontiguousarray(audio_signal, np.float32)
        B, _, T = a.shape
        ln = np.ascontiguousarray(length, np.int64)
        t_enc = T
        for _ in range(3):
            t_enc = (t_enc - 1) // 2 + 1
        out = np.zeros((B, 1024, t_enc), np.float32)
        el = np.zeros(B, np.int64)
        self._chk(self._lib.pkb_encoder_offline_step(self._e, B, T, _fptr(a), _lptr(ln), _fptr(out), _lptr(el)))
        return out, el

    def offline_utterances(self, sids, audio=None, features=None, per_feature_norm: bool = True, bins_major: bool = True,
                           want_encoder_output: bool = False, decode=True):
        """Whole-utterance offline path (pkb_offline_utterances): `audio` = list of 1-D f32 PCM arrays, or `features` = list of
        [128,T] (bins_major) / [T,128] arrays; one freshly opened stream id per utterance.  Returns the list of
        encoder_output [1024,T_enc] arrays (or None); tokens / decode trace through tokens(s) / last_steps(s)."""
        n = len(sids)
        ids = np.ascontiguousarray(sids, np.int32)
        fpp = C.POINTER(C.c_float) * n
        keep = []
        a_ptrs = ns = f_ptrs = nf = None
        t_frames = []
        if audio is not None:
            keep = [np.ascontiguousarray(a, np.float32) for a in audio]
            a_ptrs = fpp(*[_fptr(a) for a in keep])
            ns = (C.c_size_t * n)(*[a.size for a in keep])
            t_frames = [(a.size - 400) // 160 + 1 for a in keep]
        else:
            keep = [np.ascontiguousarray(f, np.float32) for f in features]
            f_ptrs = fpp(*[_fptr(f) for f in keep])
            t_frames = [f.shape[1] if bins_major else f.shape[0] for f in keep]
            nf = (C.c_int32 * n)(*t_frames)
        outs = o_ptrs = None
        if want_encoder_output:
            outs = [np.zeros((1024, self._lib.pkb_encoded_length(t)), np.float32) for t in t_frames]
            o_ptrs = fpp(*[_fptr(o) for o in outs])
        self._chk(self._lib.pkb_offline_utterances(self._e, n, ids.ctypes.data_as(C.POINTER(C.c_int32)), a_ptrs, ns, int(per_feature_norm),
                                                   f_ptrs, nf, int(bins_major), o_ptrs, int(decode)))
        return outs

    def offline_decode_pending(self) -> int:
        """Decode every utterance parked by offline_utterances(decode=2) in one batched TDT loop."""
        return self._chk(self._lib.pkb_offline_decode_pending(self._e))

    def predictor_step(self, y, h, c):
        y = np.ascontiguousarray(y, np.int64)
        h, c = np.ascontiguousarray(h, np.float32), np.ascontiguousarray(c, np.float32)
        B = y.shape[0]
        g = np.zeros((B, 640, 1), np.float32)
        ho, co = np.zeros_like(h), np.zeros_like(c)
        self._chk(self._lib.pkb_predictor_step(self._e, B, _lptr(y), _fptr(h), _fptr(c), _fptr(g), _fptr(ho), _fptr(co)))
        return g, ho, co

    def joint_step(self, enc, pred):
        enc, pred = np.ascontiguousarray(enc, np.float32), np.ascontiguousarray(pred, np.float32)
        B, _, T = enc.shape
        U = pred.shape[2]
        out = np.zeros((B, T, U, 8198), np.float32)
        self._chk(self._lib.pkb_joint_step(self._e, B, T, U, _fptr(enc), _fptr(pred), _fptr(out)))
        return out

    def logmel(self, pcm: np.ndarray, per_feature_norm: bool = False) -> np.ndarray:
        a = np.ascontiguousarray(pcm, np.float32)
        T = 0 if a.size < 400 else (a.size - 400) // 160 + 1
        out = np.zeros((max(T, 1), 128), np.float32)
        n = self._lib.pkb_logmel(self._e, _fptr(a), a.size, _fptr(out), out.size, int(per_feature_norm))
        self._chk(int(n))
        return out[:n]

    def gemm_test(self, backend: int, A: np.ndarray, W_bf16_bits: np.ndarray) -> np.ndarray:
        A = np.ascontiguousarray(A, np.float32)
        W = np.ascontiguousarray(W_bf16_bits, np.uint16)
        M, K = A.shape
        N = W.shape[0]
        Cm = np.zeros((M, N), np.float32)
        self._chk(self._lib.pkb_gemm_test(self._e, backend, M, N, K, _fptr(A), W.ctypes.data_as(C.POINTER(C.c_uint16)), _fptr(Cm)))
        return Cm
