// capi.cpp -- extern "C" boundary: parakeet_trt.h (drop-in), trt_asr.h (shim), parakeet_b200.h (additive).
//
// Conventions kept from the reference runtime (/root/reference/cpp/src/parakeet_trt.cpp):
//   create -> NULL + stderr message on failure (:1839-1842); push -> 0 / -1 / -2 with an ERROR event (:1967-1969, 3850-3857);
//   poll -> strings owned by the session until the next poll (:3860-3876); PARTIAL_TEXT at most every 100 ms of wall
//   clock when the token count changed (:3680-3712); FINAL_TEXT per push only when PARAKEET_EMIT_FINAL_EACH_CHUNK is set
//   (default off for a streaming encoder, :3802-3815); exceptions never cross the ABI.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/parakeet_b200.h"
#include "../../include/parakeet_trt.h"
#include "../../include/trt_asr.h"
#include "engine.h"

namespace {

thread_local std::string g_last_error;

bool env_bool(const char* name, bool dflt) {
  const char* v = std::getenv(name);
  if (!v || !*v) return dflt;
  return !(v[0] == '0' || v[0] == 'f' || v[0] == 'F' || v[0] == 'n' || v[0] == 'N');
}
long env_long(const char* name, long dflt) {
  const char* v = std::getenv(name);
  if (!v || !*v) return dflt;
  char* end = nullptr;
  const long r = std::strtol(v, &end, 10);
  return end == v ? dflt : r;
}
float env_float(const char* name, float dflt) {
  const char* v = std::getenv(name);
  if (!v || !*v) return dflt;
  char* end = nullptr;
  const float r = std::strtof(v, &end);
  return end == v ? dflt : r;
}

struct EventInternal {
  ParakeetEventType type;
  std::string text, err;
};

template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -2;
  } catch (...) {
    g_last_error = "unknown error";
    return -2;
  }
}

}  // namespace

// Every ABI entry runs on the engine's device, whatever device the calling thread had current (a Rust worker thread, a process
// that holds one engine per GPU), and restores the caller's device on exit.
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

struct PkbEngine {
  pkb::Engine* eng = nullptr;
  std::mutex mu;   // one caller at a time per engine
};

// One engine per (model_dir, device, options) shared by every legacy session of the process: the reference's multi-session use
// (MAGNOLIA_INTEGRATION_HANDOFF.md:10-12; SURVEY.md 8b "existing 6 symbols become a B=1 view") must not load the weights once per
// session, and pushes that arrive together are served by ONE batched pass.  Sessions are stream slots of the shared engine.
struct SharedEngine {
  pkb::Engine* eng = nullptr;
  std::string key;
  int device = 0;
  int refs = 0;
  std::mutex mu;                       // engine state + the fields below
  std::condition_variable cv_done;     // followers: "your chunk has been processed"
  std::condition_variable cv_leader;   // leader: "another session has staged a chunk"
  bool leader_active = false;
  int queued = 0;
  std::vector<ParakeetSession*> sessions;
};

struct PkbVocab {
  pkb::Vocab v;
};

struct PkbFrontend {
  pkb::Frontend* fe = nullptr;
  cudaStream_t st = nullptr;
  int device = 0, sms = 148;
  std::mutex mu;
};

struct ParakeetSession {
  SharedEngine* sh = nullptr;
  pkb::Engine* eng = nullptr;    // == sh->eng
  int sid = -1;
  bool waiting = false;          // a staged chunk awaits the next batched pass
  std::string pass_error;        // set by the leader when that pass failed
  std::mutex event_mu;
  std::queue<EventInternal> events;
  std::string last_text, last_err;
  size_t last_partial_tokens = 0;
  std::chrono::steady_clock::time_point last_partial_emit;
  uint64_t dbg_steps_left = 0;   // PARAKEET_DEBUG_TDT_STEPS: decode steps still to be traced on stderr
  bool offline = false;      // PARAKEET_B200_ENCODER=offline: the reference's non-streaming encoder engine
  bool snapshot_done = false; // PARAKEET_TDT_SNAPSHOT_DIR: step-0 tensors are dumped once per session
  std::string dbg_id;
  uint64_t dbg_utt = 0, dbg_chunk = 0, dbg_feat = 0;
};

namespace {

std::mutex g_registry_mu;
std::vector<SharedEngine*> g_registry;

// Find (or create) the shared engine of this configuration with a free stream slot and make `s` one of its streams.
void attach_shared(ParakeetSession* s, const pkb::EngineOptions& o) {
  std::ostringstream k;
  k << o.model_dir << '|' << o.device_id << '|' << o.precision << '|' << o.gemm_backend << '|' << o.punct_suppress << '|' << o.blank_penalty << '|'
    << o.max_streams;
  const std::string key = k.str();
  std::lock_guard<std::mutex> reg(g_registry_mu);
  for (SharedEngine* sh : g_registry) {
    if (sh->key != key) continue;
    std::lock_guard<std::mutex> lk(sh->mu);
    if (sh->refs >= o.max_streams) continue;      // full: look for (or create) the next engine of this configuration
    s->sid = sh->eng->open_stream();
    s->sh = sh; s->eng = sh->eng;
    sh->refs += 1;
    sh->sessions.push_back(s);
    return;
  }
  std::unique_ptr<SharedEngine> sh(new SharedEngine());
  sh->key = key;
  sh->device = o.device_id;
  sh->eng = new pkb::Engine(o);
  try {
    s->sid = sh->eng->open_stream();
  } catch (...) {
    delete sh->eng;
    throw;
  }
  s->sh = sh.get(); s->eng = sh->eng;
  sh->refs = 1;
  sh->sessions.push_back(s);
  g_registry.push_back(sh.release());
}

void detach_shared(ParakeetSession* s) {
  SharedEngine* sh = s->sh;
  if (!sh) return;
  std::lock_guard<std::mutex> reg(g_registry_mu);
  bool last = false;
  {
    DeviceGuard dg(sh->device);
    std::lock_guard<std::mutex> lk(sh->mu);
    try { if (s->sid >= 0) sh->eng->close_stream(s->sid); } catch (...) {}
    sh->sessions.erase(std::remove(sh->sessions.begin(), sh->sessions.end(), s), sh->sessions.end());
    sh->refs -= 1;
    last = sh->refs <= 0;
    if (last) { delete sh->eng; sh->eng = nullptr; }
  }
  if (last) {
    g_registry.erase(std::remove(g_registry.begin(), g_registry.end(), sh), g_registry.end());
    delete sh;
  }
  s->sh = nullptr; s->eng = nullptr; s->sid = -1;
}

}  // namespace

extern "C" {

// ================================================================================================ parakeet_trt.h
ParakeetSession* parakeet_create_session(const ParakeetConfig* config) {
  if (!config || !config->model_dir) return nullptr;
  ParakeetSession* s = nullptr;
  try {
    s = new ParakeetSession();
    pkb::EngineOptions o;
    o.model_dir = config->model_dir;
    o.device_id = config->device_id;
    o.max_streams = (int)std::min(1024L, std::max(1L, env_long("PARAKEET_B200_MAX_SESSIONS", 8)));
    o.precision = (int)env_long("PARAKEET_B200_PRECISION", config->use_fp16 ? 0 : 1);
    o.gemm_backend = (int)env_long("PARAKEET_B200_GEMM", 0);
    o.punct_suppress = env_bool("PARAKEET_DISABLE_PUNCT_SUPPRESSION", false) ? 0 : 1;
    o.blank_penalty = env_float("PARAKEET_BLANK_PENALTY", 0.0f);
    DeviceGuard dg(o.device_id);
    attach_shared(s, o);
    const char* mode = std::getenv("PARAKEET_B200_ENCODER");
    s->offline = mode && std::string(mode) == "offline";
    if (s->offline) { std::lock_guard<std::mutex> lk(s->sh->mu); s->eng->set_stream_offline(s->sid, true); }
    s->dbg_steps_left = (uint64_t)std::max(0L, env_long("PARAKEET_DEBUG_TDT_STEPS", 0));
    s->last_partial_emit = std::chrono::steady_clock::now() - std::chrono::milliseconds(1000);
    return s;
  } catch (const std::exception& e) {
    std::cerr << "[parakeet_trt] create_session failed: " << e.what() << "\n";
    g_last_error = e.what();
    if (s) { detach_shared(s); delete s; }
    return nullptr;
  }
}

void parakeet_destroy_session(ParakeetSession* session) {
  if (!session) return;
  detach_shared(session);
  delete session;
}

void parakeet_reset_utterance(ParakeetSession* session) {
  if (!session) return;
  try {
    DeviceGuard dg(session->sh->device);
    std::lock_guard<std::mutex> lk(session->sh->mu);
    session->eng->reset_stream(session->sid);
  } catch (const std::exception& e) {
    std::cerr << "[parakeet_trt] reset_utterance failed: " << e.what() << "\n";
  }
  session->last_partial_tokens = 0;
  session->last_partial_emit = std::chrono::steady_clock::now() - std::chrono::milliseconds(1000);
  std::lock_guard<std::mutex> lock(session->event_mu);
  while (!session->events.empty()) session->events.pop();
}

// PARAKEET_TDT_SNAPSHOT_DIR (parakeet_trt.cpp:2315-2400, 2612-2650, 3519-3590): once per session, on its first streaming chunk,
// the tensors around encoder step 0 and joint step (t=0, u=0) are written as raw little-endian f32 files + JSON sidecars under
// that directory, with the reference's file names, so that its triage scripts (tools/onnxruntime/compare_encoder_step0.py,
// compare_joint_step0.py) can read this library's tensors in place of the TensorRT ones.
static void write_f32_file(const std::string& path, const float* p, size_t n) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(p), (std::streamsize)(n * sizeof(float)));
}
static void write_text_file(const std::string& path, const std::string& t) { std::ofstream f(path); f << t; }

struct SnapshotPre { std::vector<float> cache_ch, cache_tm, g; int cache_len = 0; int y_id = -1; };

static void snapshot_before(ParakeetSession* s, const std::string& dir, const float* feats, size_t T, SnapshotPre* pre) {
  const int L = s->eng->n_layers();
  std::filesystem::create_directories(dir);
  pre->cache_ch.assign((size_t)L * pkb::kCacheS * pkb::kDModel, 0.f);
  pre->cache_tm.assign((size_t)L * pkb::kDModel * pkb::kTimeCtx, 0.f);
  s->eng->export_stream_state(s->sid, pre->cache_ch.data(), pre->cache_tm.data(), &pre->cache_len);
  std::vector<float> h(pkb::kPredL * pkb::kPredH), c(pkb::kPredL * pkb::kPredH);
  pre->g.assign(pkb::kPredH, 0.f);
  pre->y_id = s->eng->prime_now(s->sid);           // the predictor is primed lazily; the snapshot wants g as the joint will see it
  s->eng->get_decoder_state(s->sid, h.data(), c.data(), pre->g.data());
  write_f32_file(dir + "/features_in_trt.f32", feats, (size_t)pkb::kNMels * T);
  write_f32_file(dir + "/cache_last_channel_in_trt.f32", pre->cache_ch.data(), pre->cache_ch.size());
  write_f32_file(dir + "/cache_last_time_in_trt.f32", pre->cache_tm.data(), pre->cache_tm.size());
  std::ostringstream meta;
  meta << "{\"features_shape\":[1," << pkb::kNMels << "," << T << "],\"features_valid\":" << T << ",\"features_dtype\":\"f32\","
       << "\"cache_last_channel_shape\":[1," << L << "," << pkb::kCacheS << "," << pkb::kDModel << "],\"cache_last_time_shape\":[1," << L << ","
       << pkb::kDModel << "," << pkb::kTimeCtx << "],\"cache_last_channel_len\":" << pre->cache_len << "}";
  write_text_file(dir + "/meta_enc_trt.json", meta.str());
}

static void snapshot_after(ParakeetSession* s, const std::string& dir, const SnapshotPre& pre) {
  const pkb::ChunkResult& r = s->eng->last_chunk(s->sid);
  // cache_last_channel_len_out: raw int64 + the JSON the reference writes next to it
  const int64_t len_out = r.cache_len_out;
  { std::ofstream f(dir + "/cache_last_channel_len_out_trt.bin", std::ios::binary); f.write(reinterpret_cast<const char*>(&len_out), 8); }
  std::ostringstream lj;
  lj << "{\"dtype\":\"i64\",\"shape\":\"[1]\",\"bytes\":8,\"raw_i32\":" << (int32_t)len_out << ",\"raw_i64\":" << len_out << ",\"raw_f32\":\"nan\",\"raw\":"
     << len_out << ",\"effective\":" << len_out << "}";
  write_text_file(dir + "/cache_last_channel_len_out_trt.json", lj.str());
  // encoder slice fed to the joint [1,1024,3], its frame 0, the predictor output of step (0,0), the duration logits of that step
  std::vector<float> enc((size_t)pkb::kDModel * pkb::kValidOut, 0.f), enc_t0(pkb::kDModel), logits(pkb::kJointOut);
  const int T_enc = s->eng->last_encoder_output(s->sid, enc.data(), pkb::kValidOut);
  const int Tc = std::min(T_enc, (int)pkb::kValidOut);
  for (int c = 0; c < pkb::kDModel; ++c) enc_t0[c] = Tc > 0 ? enc[(size_t)c * Tc] : 0.f;
  s->eng->joint_step(1, 1, 1, enc_t0.data(), pre.g.data(), logits.data());
  write_f32_file(dir + "/enc_slice_trt.f32", enc.data(), (size_t)pkb::kDModel * Tc);
  write_f32_file(dir + "/enc_out_t0_trt.f32", enc_t0.data(), enc_t0.size());
  write_f32_file(dir + "/pred_g_trt.f32", pre.g.data(), pre.g.size());
  write_f32_file(dir + "/dur_logits_trt.f32", logits.data() + pkb::kVocab, pkb::kNDur);
  const int best_tok = r.steps.empty() ? -1 : r.steps[0].token, best_dur = r.steps.empty() ? -1 : r.steps[0].duration;
  std::ostringstream meta;
  meta << "{\"enc_shape\":[1," << pkb::kDModel << "," << Tc << "],\"enc_out_t0_shape\":[1," << pkb::kDModel << ",1],\"pred_shape\":[1," << pkb::kPredH
       << ",1],\"dur_shape\":[" << pkb::kNDur << "],\"tok_offset\":0,\"dur_offset\":" << pkb::kVocab << ",\"token_span\":" << pkb::kVocab
       << ",\"dur_bins_used\":" << pkb::kNDur << ",\"best_tok\":" << best_tok << ",\"best_dur_idx\":" << best_dur << ",\"y_id\":" << pre.y_id << "}";
  write_text_file(dir + "/meta_trt.json", meta.str());
  std::cerr << "[parakeet_trt] tdt_snapshot dir=" << dir << " enc=" << dir << "/enc_slice_trt.f32 pred=" << dir << "/pred_g_trt.f32 dur=" << dir
            << "/dur_logits_trt.f32 enc_out_t0=" << dir << "/enc_out_t0_trt.f32\n";
}

// NaN / Inf guard after the encoder step (parakeet_trt.cpp:913-1013, 2449-2481): the first 10 guarded tensors of the process, then one in
// 100, or every one with PARAKEET_NAN_GUARD_ALWAYS=1; findings go to stderr in the reference's line format; PARAKEET_NAN_GUARD_HALT=1
// aborts the process on the first finding.
static void nan_guard_after_chunk(ParakeetSession* s, int cache_len_in, size_t T) {
  static std::atomic<int> s_guard_count{0};
  static const char* kStage[3] = {"enc_output", "enc_cache_ch_out", "enc_cache_tm_out"};
  const bool force = env_bool("PARAKEET_NAN_GUARD_ALWAYS", false);
  for (int stage = 0; stage < (s->offline ? 1 : 3); ++stage) {
    const int n = s_guard_count.fetch_add(1, std::memory_order_relaxed);
    if (!force && n >= 10 && (n % 100) != 0) continue;
    const pkb::Engine::GuardResult r = s->eng->nan_guard(s->sid, stage);
    if (r.nan_count == 0 && r.inf_count == 0) continue;
    std::cerr << "[parakeet_trt] NAN_GUARD ALERT stage=" << kStage[stage] << " nan_count=" << r.nan_count << " inf_count=" << r.inf_count
              << " finite_count=" << (r.sample_n - (size_t)r.nan_count - (size_t)r.inf_count) << " sample_n=" << r.sample_n << "/" << r.count
              << " dtype=" << (stage == 1 && s->eng->options().precision == 0 ? "bf16" : "fp32") << " cache_len_in=" << cache_len_in
              << " length_in=" << T;
    if (stage == 0) std::cerr << " chunk_idx=" << s->dbg_chunk << " feature_idx=" << s->dbg_feat;
    if (r.nan_count > 0) std::cerr << " first_nan_idx=" << r.first_nan;
    if (r.inf_count > 0) std::cerr << " first_inf_idx=" << r.first_inf;
    std::cerr << "\n";
    if (env_bool("PARAKEET_NAN_GUARD_HALT", false)) {
      std::cerr << "[parakeet_trt] NAN_GUARD_HALT enabled, aborting\n";
      std::abort();
    }
  }
}

// Everything the reference does after the decode loop of one push (trace lines, partial / final events), for one session whose chunk
// has just been processed.  Called with the shared engine locked.
static void after_chunk(ParakeetSession* s, size_t tokens_before) {
  const std::vector<int>& toks = s->eng->tokens(s->sid);
  if (s->dbg_steps_left > 0) {
    // decode trace in the line format the reference prints and tools/verify_nemo/compare_tdt_trace.py:43-66 parses
    // (--cpp-stderr): the first PARAKEET_DEBUG_TDT_STEPS steps of the session (parakeet_trt.cpp:2874)
    int prev_t = -1, u = 0;
    for (const pkb::StepRecord& r : s->eng->last_chunk(s->sid).steps) {
      if (s->dbg_steps_left == 0) break;
      u = r.time_idx == prev_t ? u + 1 : 0;
      prev_t = r.time_idx;
      const bool blank = r.token == pkb::kBlank, clamped = blank && r.duration == 0;
      std::cerr << "[parakeet_trt] tdt_step time_idx=" << r.time_idx << " u=" << u << " best_tok=" << r.token << " best_dur_idx=" << r.duration
                << " duration=" << r.duration << " advance=" << (clamped ? 1 : r.duration) << " blank=" << (blank ? 1 : 0)
                << " blank_dur0_clamped=" << (clamped ? 1 : 0) << "\n";
      --s->dbg_steps_left;
    }
  }
  const auto now = std::chrono::steady_clock::now();
  if (now - s->last_partial_emit >= std::chrono::milliseconds(100)) {
    if (toks.size() != s->last_partial_tokens) {
      s->last_partial_tokens = toks.size();
      EventInternal ev{PARAKEET_EVENT_PARTIAL_TEXT, s->eng->detokenize(toks), ""};
      std::lock_guard<std::mutex> lock(s->event_mu);
      s->events.push(std::move(ev));
    }
    s->last_partial_emit = now;
  }
  if (env_bool("PARAKEET_EMIT_FINAL_EACH_CHUNK", s->offline)) {      // default: !enc_streaming (parakeet_trt.cpp:3802)
    std::vector<int> chunk(toks.begin() + (std::ptrdiff_t)tokens_before, toks.end());
    EventInternal ev{PARAKEET_EVENT_FINAL_TEXT, s->eng->detokenize(chunk), ""};
    std::lock_guard<std::mutex> lock(s->event_mu);
    s->events.push(std::move(ev));
  }
}

// One push = one encoder chunk of this session.  The chunk is staged in the shared engine; the first session to arrive becomes the
// leader of the next batched pass, waits a moment for the other sessions that are pushing right now (PARAKEET_B200_COALESCE_US, default
// 200 us, only when the engine has more than one session), runs ONE pass for every staged chunk and wakes the followers.
static void push_one_chunk(ParakeetSession* s, const float* feats, size_t T) {
  SharedEngine* sh = s->sh;
  DeviceGuard dg(sh->device);
  std::unique_lock<std::mutex> lk(sh->mu);
  const size_t before = s->eng->tokens(s->sid).size();
  const int cache_len_in = s->eng->cache_len(s->sid);
  const char* snap = std::getenv("PARAKEET_TDT_SNAPSHOT_DIR");
  const bool want_snap = snap && *snap && !s->offline && !s->snapshot_done;
  SnapshotPre pre;
  if (want_snap) {
    try { snapshot_before(s, snap, feats, T, &pre); }
    catch (const std::exception& e) { std::cerr << "[parakeet_trt] WARN: failed to write encoder snapshot: " << e.what() << "\n"; }
  }
  s->eng->queue_features(s->sid, feats, (int)T);      // staging only; throws on a chunk the encoder cannot take (nothing is queued then)
  s->waiting = true;
  s->pass_error.clear();
  sh->queued += 1;
  if (!sh->leader_active) {
    sh->leader_active = true;
    if (sh->refs > 1) {
      static const long window_us = std::max(0L, env_long("PARAKEET_B200_COALESCE_US", 200));
      if (window_us > 0) sh->cv_leader.wait_for(lk, std::chrono::microseconds(window_us), [&] { return sh->queued >= sh->refs; });
    }
    std::string err;
    try {
      s->eng->step();
    } catch (const std::exception& e) {
      err = e.what();
    }
    for (ParakeetSession* o : sh->sessions) {
      if (!o->waiting) continue;
      o->waiting = false;
      o->pass_error = err;
      if (!err.empty()) { try { sh->eng->drop_pending(o->sid); } catch (...) {} }
    }
    sh->queued = 0;
    sh->leader_active = false;
    sh->cv_done.notify_all();
  } else {
    sh->cv_leader.notify_one();
    sh->cv_done.wait(lk, [&] { return !s->waiting; });
  }
  if (!s->pass_error.empty()) throw std::runtime_error(s->pass_error);
  nan_guard_after_chunk(s, cache_len_in, T);
  if (want_snap) {
    try { snapshot_after(s, snap, pre); }
    catch (const std::exception& e) { std::cerr << "[parakeet_trt] WARN: failed to write TDT snapshot: " << e.what() << "\n"; }
    s->snapshot_done = true;
  }
  after_chunk(s, before);
}

int parakeet_push_features(ParakeetSession* session, const float* features, size_t num_frames) {
  if (!session || !features) return -1;
  if (num_frames == 0) return 0;
  try {
    long max_frames = env_long("PARAKEET_MAX_FRAMES_PER_PUSH", 256);
    if (max_frames < 1 || max_frames > 256) max_frames = 256;
    if (num_frames <= (size_t)max_frames) {
      push_one_chunk(session, features, num_frames);
      return 0;
    }
    // auto-chunk: re-slice the [128, num_frames] rows (parakeet_trt.cpp:1989-2011)
    std::vector<float> buf((size_t)pkb::kNMels * (size_t)max_frames);
    for (size_t off = 0; off < num_frames; off += (size_t)max_frames) {
      const size_t n = std::min((size_t)max_frames, num_frames - off);
      for (int m = 0; m < pkb::kNMels; ++m)
        std::memcpy(buf.data() + (size_t)m * n, features + (size_t)m * num_frames + off, n * sizeof(float));
      push_one_chunk(session, buf.data(), n);
    }
    return 0;
  } catch (const std::exception& e) {
    EventInternal ev{PARAKEET_EVENT_ERROR, "", e.what()};
    std::lock_guard<std::mutex> lock(session->event_mu);
    session->events.push(std::move(ev));
    return -2;
  }
}

void parakeet_set_debug_context(ParakeetSession* session, const char* id, uint64_t utt_seq, uint64_t audio_chunk_idx,
                                uint64_t feature_idx) {
  if (!session) return;
  if (id) session->dbg_id = id; else session->dbg_id.clear();
  session->dbg_utt = utt_seq;
  session->dbg_chunk = audio_chunk_idx;
  session->dbg_feat = feature_idx;
}

bool parakeet_poll_event(ParakeetSession* session, ParakeetEvent* event) {
  if (!session || !event) return false;
  std::lock_guard<std::mutex> lock(session->event_mu);
  if (session->events.empty()) return false;
  EventInternal& ev = session->events.front();
  session->last_text = ev.text;
  session->last_err = ev.err;
  event->type = ev.type;
  event->segment_id = 0;
  event->text = session->last_text.c_str();
  event->error_message = session->last_err.c_str();
  session->events.pop();
  return true;
}

// ================================================================================================ trt_asr.h
}  // extern "C"

struct TrtAsrSession {
  ParakeetSession* inner = nullptr;
  std::string text, err;
};

namespace {
float half_to_float(uint16_t u) {   // IEEE binary16 -> binary32; subnormals flush to signed zero like the reference shim (trt_asr.cpp:21-37)
  const uint32_t sign = (uint32_t)(u & 0x8000u) << 16, e = (u >> 10) & 0x1Fu, m = u & 0x3FFu;
  uint32_t bits;
  if (e == 0) bits = sign;
  else if (e == 31) bits = sign | 0x7F800000u | (m << 13);
  else bits = sign | ((e + 112u) << 23) | (m << 13);
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}
}  // namespace

extern "C" {

TrtAsrSession* trt_asr_create_session(const TrtAsrConfig* config) {
  if (!config || !config->model_dir) return nullptr;
  ParakeetConfig pc{config->model_dir, config->device_id, config->use_fp16};
  ParakeetSession* inner = parakeet_create_session(&pc);
  if (!inner) return nullptr;
  TrtAsrSession* s = new TrtAsrSession();
  s->inner = inner;
  return s;
}
void trt_asr_destroy_session(TrtAsrSession* session) {
  if (!session) return;
  parakeet_destroy_session(session->inner);
  delete session;
}
void trt_asr_reset_session(TrtAsrSession* session) {
  if (!session) return;
  parakeet_reset_utterance(session->inner);
  session->text.clear();
  session->err.clear();
}
int trt_asr_push_features_f32(TrtAsrSession* session, const float* features_f32, int32_t T, int32_t length) {
  (void)length;
  if (!session || !features_f32 || T <= 0) return -1;
  return parakeet_push_features(session->inner, features_f32, (size_t)T);
}
int trt_asr_push_features_f16(TrtAsrSession* session, const uint16_t* features_f16, int32_t T, int32_t length) {
  if (!session || !features_f16 || T <= 0) return -1;
  std::vector<float> tmp((size_t)pkb::kNMels * (size_t)T);
  for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = half_to_float(features_f16[i]);
  return trt_asr_push_features_f32(session, tmp.data(), T, length);
}
bool trt_asr_poll_event(TrtAsrSession* session, TrtAsrEvent* out_event) {
  if (!session || !out_event) return false;
  ParakeetEvent ev{};
  if (!parakeet_poll_event(session->inner, &ev)) return false;
  out_event->segment_id = ev.segment_id;
  out_event->token_id = -1;
  out_event->text = nullptr;
  out_event->error_message = nullptr;
  if (ev.type == PARAKEET_EVENT_PARTIAL_TEXT || ev.type == PARAKEET_EVENT_FINAL_TEXT) {
    out_event->type = ev.type == PARAKEET_EVENT_PARTIAL_TEXT ? TRT_ASR_EVENT_PARTIAL_TEXT : TRT_ASR_EVENT_FINAL_TEXT;
    session->text = ev.text ? ev.text : "";
    out_event->text = session->text.c_str();
  } else {
    out_event->type = TRT_ASR_EVENT_ERROR;
    session->err = (ev.error_message && *ev.error_message) ? ev.error_message : "unknown error";
    out_event->error_message = session->err.c_str();
  }
  return true;
}

// ================================================================================================ parakeet_b200.h
const char* pkb_last_error(void) { return g_last_error.c_str(); }
const char* pkb_version(void) { return "parakeet-b200 0.1 (sm_100a)"; }

PkbEngine* pkb_engine_create(const PkbEngineConfig* c) {
  if (!c || !c->model_dir) { g_last_error = "null config"; return nullptr; }
  try {
    pkb::EngineOptions o;
    o.model_dir = c->model_dir;
    o.device_id = c->device_id;
    o.max_streams = c->max_streams > 0 ? c->max_streams : 1;
    o.precision = c->precision;
    o.gemm_backend = c->gemm_backend;
    o.contract_cache = c->contract_cache;
    o.punct_suppress = c->punct_suppression;
    o.max_rows = c->max_rows;
    o.blank_penalty = env_float("PARAKEET_BLANK_PENALTY", 0.0f);
    DeviceGuard dg(o.device_id);
    PkbEngine* e = new PkbEngine();
    e->eng = new pkb::Engine(o);
    return e;
  } catch (const std::exception& ex) {
    g_last_error = ex.what();
    return nullptr;
  }
}
void pkb_engine_destroy(PkbEngine* e) {
  if (!e) return;
  DeviceGuard dg(e->eng->options().device_id);
  delete e->eng;
  delete e;
}
int32_t pkb_engine_num_layers(PkbEngine* e) { return e ? e->eng->n_layers() : -1; }
int64_t pkb_engine_kernel_launches(PkbEngine* e) { return e ? e->eng->kernel_launches() : -1; }

#define PKB_ENTER(e)                                        \
  if (!(e)) { g_last_error = "null engine"; return -1; }    \
  DeviceGuard _dg((e)->eng->options().device_id);           \
  std::lock_guard<std::mutex> _lock((e)->mu)

int32_t pkb_stream_open(PkbEngine* e) { PKB_ENTER(e); return guarded([&] { return e->eng->open_stream(); }); }
int32_t pkb_stream_close(PkbEngine* e, int32_t s) { PKB_ENTER(e); return guarded([&] { e->eng->close_stream(s); return 0; }); }
int32_t pkb_stream_reset(PkbEngine* e, int32_t s) { PKB_ENTER(e); return guarded([&] { e->eng->reset_stream(s); return 0; }); }
int32_t pkb_stream_push_features(PkbEngine* e, int32_t s, const float* f, int32_t T) {
  PKB_ENTER(e);
  if (!f) { g_last_error = "null features"; return -1; }
  return guarded([&] { e->eng->queue_features(s, f, T); return 0; });
}
int32_t pkb_stream_push_audio(PkbEngine* e, int32_t s, const float* pcm, size_t n) {
  PKB_ENTER(e);
  if (!pcm && n) { g_last_error = "null pcm"; return -1; }
  return guarded([&] { e->eng->queue_audio(s, pcm, n); return 0; });
}
int32_t pkb_stream_set_feature_norm(PkbEngine* e, int32_t s, const float* mean128, const float* std128) {
  PKB_ENTER(e);
  return guarded([&] { e->eng->set_feature_norm(s, mean128, std128); return 0; });
}
int32_t pkb_stream_set_feature_norm_running(PkbEngine* e, int32_t s, int32_t on) {
  PKB_ENTER(e);
  return guarded([&] { e->eng->set_feature_norm_running(s, on != 0); return 0; });
}
int32_t pkb_stream_set_offline(PkbEngine* e, int32_t s, int32_t offline) {
  PKB_ENTER(e);
  return guarded([&] { e->eng->set_stream_offline(s, offline != 0); return 0; });
}
int32_t pkb_engine_push_audio_batch(PkbEngine* e, int32_t n, const int32_t* sids, const float* pcm, int64_t stride, int32_t count) {
  PKB_ENTER(e);
  if (!sids || (!pcm && n > 0 && count > 0)) { g_last_error = "null argument"; return -1; }
  return guarded([&] { e->eng->push_audio_batch(n, sids, pcm, stride, count, false); return 0; });
}
int32_t pkb_engine_push_audio_batch_device(PkbEngine* e, int32_t n, const int32_t* sids, const float* d_pcm, int64_t stride,
                                           int32_t count) {
  PKB_ENTER(e);
  if (!sids || (!d_pcm && n > 0 && count > 0)) { g_last_error = "null argument"; return -1; }
  return guarded([&] { e->eng->push_audio_batch(n, sids, d_pcm, stride, count, true); return 0; });
}
int32_t pkb_engine_event_record(PkbEngine* e) { PKB_ENTER(e); return guarded([&] { return e->eng->event_record(); }); }
double pkb_engine_event_elapsed_ms(PkbEngine* e, int32_t a, int32_t b) {
  if (!e) return -1.0;
  DeviceGuard dg(e->eng->options().device_id);
  std::lock_guard<std::mutex> lock(e->mu);
  try { return e->eng->event_elapsed_ms(a, b); } catch (const std::exception& ex) { g_last_error = ex.what(); return -1.0; }
}
int32_t pkb_engine_profile_enable(PkbEngine* e, int32_t on) { PKB_ENTER(e); return guarded([&] { e->eng->profile_enable(on != 0); return 0; }); }
int32_t pkb_engine_profile_read(PkbEngine* e, double* ms, double* flops, int64_t* launches) {
  PKB_ENTER(e);
  return guarded([&] { long long l = 0; e->eng->profile_read(0, ms, flops, &l); *launches = l; return 0; });
}
int32_t pkb_engine_profile_read_class(PkbEngine* e, int32_t cls, double* ms, double* work, int64_t* launches) {
  PKB_ENTER(e);
  return guarded([&] { long long l = 0; e->eng->profile_read(cls, ms, work, &l); *launches = l; return 0; });
}
int32_t pkb_engine_decode_loop_stats(PkbEngine* e, double* ms, double* bytes, int64_t* passes, int64_t* loops, int32_t reset) {
  PKB_ENTER(e);
  if (!ms || !bytes || !passes || !loops) { g_last_error = "null argument"; return -1; }
  return guarded([&] { long long p = 0, l = 0; e->eng->decode_loop_stats(ms, bytes, &p, &l, reset); *passes = p; *loops = l; return 0; });
}
int32_t pkb_engine_set_blank_penalty(PkbEngine* e, float penalty) {
  PKB_ENTER(e);
  return guarded([&] { e->eng->set_blank_penalty(penalty); return 0; });
}
int32_t pkb_engine_graphs_built(PkbEngine* e) { PKB_ENTER(e); return e->eng->graphs_built(); }
int32_t pkb_engine_step(PkbEngine* e) { PKB_ENTER(e); return guarded([&] { return e->eng->step(); }); }
int32_t pkb_stream_has_pending(PkbEngine* e, int32_t s) { PKB_ENTER(e); return guarded([&] { return e->eng->has_pending(s) ? 1 : 0; }); }
int32_t pkb_stream_num_tokens(PkbEngine* e, int32_t s) { PKB_ENTER(e); return guarded([&] { return (int)e->eng->tokens(s).size(); }); }
int32_t pkb_stream_tokens(PkbEngine* e, int32_t s, int32_t* out, int32_t cap) {
  PKB_ENTER(e);
  return guarded([&] {
    const std::vector<int>& t = e->eng->tokens(s);
    for (int i = 0; i < cap && i < (int)t.size(); ++i) out[i] = t[i];
    return (int)t.size();
  });
}
int32_t pkb_stream_last_steps(PkbEngine* e, int32_t s, PkbStep* out, int32_t cap) {
  PKB_ENTER(e);
  return guarded([&] {
    const pkb::ChunkResult& r = e->eng->last_chunk(s);
    for (int i = 0; i < cap && i < (int)r.steps.size(); ++i) out[i] = PkbStep{r.steps[i].time_idx, r.steps[i].token, r.steps[i].duration};
    return (int)r.steps.size();
  });
}
int32_t pkb_stream_token_frames(PkbEngine* e, int32_t s, int32_t* out, int32_t cap) {
  PKB_ENTER(e);
  return guarded([&] {
    const std::vector<int>& t = e->eng->token_frames(s);
    for (int i = 0; i < cap && i < (int)t.size(); ++i) out[i] = t[i];
    return (int)t.size();
  });
}
int64_t pkb_stream_encoder_frames(PkbEngine* e, int32_t s) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lock(e->mu);
  try { return e->eng->encoder_frames_done(s); } catch (const std::exception& ex) { g_last_error = ex.what(); return -2; }
}
int32_t pkb_stream_stable_prefix(PkbEngine* e, int32_t s, int32_t revision_window_ms) {
  PKB_ENTER(e);
  return guarded([&] { return e->eng->stable_prefix(s, revision_window_ms); });
}
int32_t pkb_stream_cache_len(PkbEngine* e, int32_t s) { PKB_ENTER(e); return guarded([&] { return e->eng->cache_len(s); }); }
int64_t pkb_stream_chunks_done(PkbEngine* e, int32_t s) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lock(e->mu);
  try { return e->eng->chunks_done(s); } catch (const std::exception& ex) { g_last_error = ex.what(); return -2; }
}
static int32_t copy_text(const std::string& t, char* out, int32_t cap) {
  if (out && cap > 0) {
    const size_t n = std::min((size_t)cap - 1, t.size());
    std::memcpy(out, t.data(), n);
    out[n] = 0;
  }
  return (int32_t)t.size();
}
int32_t pkb_stream_text(PkbEngine* e, int32_t s, char* out, int32_t cap) {
  PKB_ENTER(e);
  return guarded([&] { return copy_text(e->eng->detokenize(e->eng->tokens(s)), out, cap); });
}
int32_t pkb_detokenize(PkbEngine* e, const int32_t* ids, int32_t n, char* out, int32_t cap) {
  PKB_ENTER(e);
  return guarded([&] { return copy_text(e->eng->detokenize(std::vector<int>(ids, ids + n)), out, cap); });
}
int32_t pkb_token_is_punct_only(PkbEngine* e, int32_t id) {
  PKB_ENTER(e);
  return e->eng->vocab().is_punct_only(id) ? 1 : 0;
}
// ---- the token table on its own (no GPU): what cpp/src/tokenizer.cpp gives the reference's callers
PkbVocab* pkb_vocab_open(const char* vocab_path) {
  if (!vocab_path) { g_last_error = "null path"; return nullptr; }
  try {
    PkbVocab* v = new PkbVocab();
    v->v = pkb::Vocab(vocab_path);
    return v;
  } catch (const std::exception& ex) {
    g_last_error = ex.what();
    return nullptr;
  }
}
void pkb_vocab_close(PkbVocab* v) { delete v; }
int32_t pkb_vocab_size(const PkbVocab* v) { return v ? v->v.size() : -1; }
int32_t pkb_vocab_decode(const PkbVocab* v, const int32_t* ids, int32_t n, char* out, int32_t cap) {
  if (!v || (!ids && n > 0) || n < 0) { g_last_error = "null argument"; return -1; }
  return copy_text(v->v.decode(ids, (size_t)n), out, cap);
}
int32_t pkb_vocab_is_punct_only(const PkbVocab* v, int32_t id) { return v && v->v.is_punct_only(id) ? 1 : 0; }
int32_t pkb_stream_import_state(PkbEngine* e, int32_t s, const float* cc, const float* ct, int32_t len) {
  PKB_ENTER(e);
  if (!cc || !ct) { g_last_error = "null argument"; return -1; }
  return guarded([&] { e->eng->import_stream_state(s, cc, ct, len); return 0; });
}
int32_t pkb_stream_export_state(PkbEngine* e, int32_t s, float* cc, float* ct, int32_t* len) {
  PKB_ENTER(e);
  if (!cc || !ct || !len) { g_last_error = "null argument"; return -1; }
  return guarded([&] { int l = 0; e->eng->export_stream_state(s, cc, ct, &l); *len = l; return 0; });
}
int32_t pkb_stream_set_decoder_state(PkbEngine* e, int32_t s, const float* h, const float* c, const float* g, int32_t n_emitted, int32_t y_id) {
  PKB_ENTER(e);
  if (!h || !c || !g) { g_last_error = "null argument"; return -1; }
  return guarded([&] { e->eng->set_decoder_state(s, h, c, g, n_emitted, y_id); return 0; });
}
int32_t pkb_stream_get_decoder_state(PkbEngine* e, int32_t s, float* h, float* c, float* g) {
  PKB_ENTER(e);
  if (!h || !c || !g) { g_last_error = "null argument"; return -1; }
  return guarded([&] { e->eng->get_decoder_state(s, h, c, g); return 0; });
}
int32_t pkb_encoder_streaming_step(PkbEngine* e, int32_t B, int32_t T, const float* audio_signal, const int64_t* length,
                                   const float* cc, const float* ct, const int64_t* cl, float* enc_out, int64_t* enc_len, float* cc_out,
                                   float* ct_out, int64_t* cl_out) {
  PKB_ENTER(e);
  return guarded([&] { e->eng->encoder_streaming_step(B, T, audio_signal, length, cc, ct, cl, enc_out, enc_len, cc_out, ct_out, cl_out); return 0; });
}
int32_t pkb_encoder_offline_step(PkbEngine* e, int32_t B, int32_t T, const float* audio_signal, const int64_t* length, float* enc_out,
                                 int64_t* enc_len) {
  PKB_ENTER(e);
  if (!audio_signal || !length || !enc_out || !enc_len) { g_last_error = "null argument"; return -1; }
  return guarded([&] { e->eng->encoder_offline_step(B, T, audio_signal, length, enc_out, enc_len); return 0; });
}
int32_t pkb_offline_utterances(PkbEngine* e, int32_t n, const int32_t* streams, const float* const* audio, const size_t* n_samples,
                               int32_t per_feature_norm, const float* const* features, const int32_t* n_frames, int32_t bins_major,
                               float* const* encoder_output, int32_t decode) {
  PKB_ENTER(e);
  if (!streams || (!audio && !features)) { g_last_error = "null argument"; return -1; }
  return guarded([&] {
    e->eng->offline_utterances(n, streams, audio, n_samples, per_feature_norm, features, n_frames, bins_major, encoder_output, decode);
    return 0;
  });
}
int32_t pkb_offline_decode_pending(PkbEngine* e) {
  PKB_ENTER(e);
  return guarded([&] { return e->eng->offline_decode_pending(); });
}
uint64_t pkb_debug_ring_valid_groups(int32_t head, int32_t len, int32_t qlen) {
  if (head < 0 || head >= pkb::kRingCap || len < 0 || len > pkb::kCacheS || qlen < 0 || qlen > pkb::kMaxTq) return 0;
  return pkb::ring_valid_groups(head, len, qlen);
}
int32_t pkb_encoded_length(int32_t L) {
  for (int i = 0; i < 3; ++i) L = L <= 0 ? 0 : (L - 1) / 2 + 1;
  return L;
}
int32_t pkb_predictor_step(PkbEngine* e, int32_t B, const int64_t* y, const float* h, const float* c, float* g, float* h_out, float* c_out) {
  PKB_ENTER(e);
  return guarded([&] { e->eng->predictor_step(B, y, h, c, g, h_out, c_out); return 0; });
}
int32_t pkb_joint_step(PkbEngine* e, int32_t B, int32_t T, int32_t U, const float* enc, const float* pred, float* out) {
  PKB_ENTER(e);
  return guarded([&] { e->eng->joint_step(B, T, U, enc, pred, out); return 0; });
}
int64_t pkb_logmel(PkbEngine* e, const float* pcm, size_t n, float* out, size_t out_cap_floats, int32_t per_feature_norm) {
  if (!e) { g_last_error = "null engine"; return -1; }
  DeviceGuard dg(e->eng->options().device_id);
  std::lock_guard<std::mutex> lock(e->mu);
  try {
    const size_t T = n < 400 ? 0 : (n - 400) / 160 + 1;
    if (T * pkb::kNMels > out_cap_floats) { g_last_error = "output buffer too small"; return -1; }
    return (int64_t)e->eng->logmel(pcm, n, out, per_feature_norm);
  } catch (const std::exception& ex) {
    g_last_error = ex.what();
    return -2;
  }
}
PkbFrontend* pkb_frontend_create(int32_t device_id) {
  try {
    DeviceGuard dg(device_id);
    cudaDeviceProp prop;
    PKB_CUDA(cudaGetDeviceProperties(&prop, device_id));
    PKB_CHECK(prop.major == 10, "this library is built for sm_100a (B200) only");
    PkbFrontend* f = new PkbFrontend();
    f->device = device_id;
    f->sms = prop.multiProcessorCount;
    f->fe = new pkb::Frontend();
    PKB_CUDA(cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking));
    PKB_CUDA(cudaDeviceSynchronize());
    return f;
  } catch (const std::exception& ex) {
    g_last_error = ex.what();
    return nullptr;
  }
}
void pkb_frontend_destroy(PkbFrontend* f) {
  if (!f) return;
  DeviceGuard dg(f->device);
  cudaStreamSynchronize(f->st);
  cudaStreamDestroy(f->st);
  delete f->fe;
  delete f;
}
int64_t pkb_frontend_logmel(PkbFrontend* f, const float* pcm, size_t n, float* out, size_t cap) {
  if (!f || !pcm || !out) { g_last_error = "null argument"; return -1; }
  std::lock_guard<std::mutex> lock(f->mu);
  float *d_audio = nullptr, *d_out = nullptr;
  pkb::FrontSegment* d_seg = nullptr;
  int* d_prefix = nullptr;
  try {
    const size_t T = n < 400 ? 0 : (n - 400) / 160 + 1;
    if (T == 0) return 0;
    if (T * pkb::kNMels > cap) { g_last_error = "output buffer too small"; return -1; }
    DeviceGuard dg(f->device);
    PKB_CUDA(cudaMalloc(&d_audio, (n + 2) * 4));
    PKB_CUDA(cudaMalloc(&d_out, T * pkb::kNMels * 4));
    PKB_CUDA(cudaMalloc(&d_seg, sizeof(pkb::FrontSegment)));
    PKB_CUDA(cudaMalloc(&d_prefix, 2 * sizeof(int)));
    const pkb::FrontSegment sg{0, 0, pkb::kNMels, 0, 0, -1};
    const int prefix[2] = {0, (int)T};
    PKB_CUDA(cudaMemcpyAsync(d_audio, pcm, n * 4, cudaMemcpyHostToDevice, f->st));
    PKB_CUDA(cudaMemcpyAsync(d_seg, &sg, sizeof(sg), cudaMemcpyHostToDevice, f->st));
    PKB_CUDA(cudaMemcpyAsync(d_prefix, prefix, sizeof(prefix), cudaMemcpyHostToDevice, f->st));
    f->fe->logmel(d_audio, d_seg, d_prefix, 1, (int)T, d_out, nullptr, f->sms, f->st);
    PKB_CUDA(cudaMemcpyAsync(out, d_out, T * pkb::kNMels * 4, cudaMemcpyDeviceToHost, f->st));
    PKB_CUDA(cudaStreamSynchronize(f->st));
    cudaFree(d_audio); cudaFree(d_out); cudaFree(d_seg); cudaFree(d_prefix);
    return (int64_t)T;
  } catch (const std::exception& ex) {
    g_last_error = ex.what();
    cudaFree(d_audio); cudaFree(d_out); cudaFree(d_seg); cudaFree(d_prefix);
    return -2;
  }
}
int32_t pkb_gemm_test(PkbEngine* e, int32_t backend, int32_t M, int32_t N, int32_t K, const float* A, const uint16_t* W, float* C) {
  PKB_ENTER(e);
  return guarded([&] { e->eng->gemm_test(backend, M, N, K, A, W, C, 0); return 0; });
}

}  // extern "C"
