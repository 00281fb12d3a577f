// parakeet_cli -- replay / stream-simulation driver over the libparakeet_trt C ABI (B200 build).
//
// C++ counterpart of the reference's Rust CLI (/root/reference/rust/cli/src/main.rs:12-47 arguments, :187-540 flow), written
// against the public headers only (include/parakeet_trt.h for push / poll, include/parakeet_b200.h for the GPU log-mel
// frontend that replaces rust/features).  Inputs: WAV (PCM16 or float32, mono, 16 kHz), raw PCM (f32le), or a feature tap
// (f32le raw + optional JSON sidecar: kind / format / layout / mel_bins / num_frames / shape, main.rs:132-165).
//   parakeet_cli <input> --model-dir DIR [--stream-sim SEC] [--device-id N] [--raw-pcm] [--sample-rate HZ] [--features-input]
//                [--n-mels N] [--verbose|-v] [--dump-features PATH] [--feature-norm none|per_feature] [--no-sleep] [--whole-utterance] [--stream-audio SECONDS]
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/parakeet_b200.h"
#include "../../include/parakeet_trt.h"

namespace {

constexpr int kMels = 128;

struct Args {
  std::string input, model_dir, dump_features, feature_norm;
  double stream_sim = -1.0;
  int device_id = 0, n_mels = -1;
  long sample_rate = -1;
  bool raw_pcm = false, features_input = false, verbose = false, no_sleep = false, whole_utterance = false;
  double stream_audio = -1.0;      // additive: cache-aware streaming from audio (pkb_stream_push_audio), seconds per push
};

[[noreturn]] void die(const std::string& msg) {
  std::fprintf(stderr, "Error: %s\n", msg.c_str());
  std::exit(1);
}

std::vector<uint8_t> read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) die("cannot open " + path);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

std::vector<float> load_f32le(const std::string& path, const char* what) {
  const std::vector<uint8_t> b = read_file(path);
  if (b.size() % 4 != 0) die(std::string(what) + " file size must be multiple of 4 (f32le)");
  std::vector<float> v(b.size() / 4);
  std::memcpy(v.data(), b.data(), b.size());
  return v;
}

// minimal RIFF/WAVE reader: PCM16 or IEEE float32, first channel only
std::vector<float> load_wav(const std::string& path, long* rate) {
  const std::vector<uint8_t> b = read_file(path);
  if (b.size() < 44 || std::memcmp(b.data(), "RIFF", 4) != 0 || std::memcmp(b.data() + 8, "WAVE", 4) != 0) die("not a RIFF/WAVE file: " + path);
  size_t pos = 12;
  int fmt = 0, channels = 1, bits = 16;
  std::vector<float> out;
  while (pos + 8 <= b.size()) {
    const std::string id((const char*)b.data() + pos, 4);
    uint32_t len;
    std::memcpy(&len, b.data() + pos + 4, 4);
    const uint8_t* p = b.data() + pos + 8;
    if (pos + 8 + len > b.size()) len = (uint32_t)(b.size() - pos - 8);
    if (id == "fmt ") {
      uint16_t f16, ch, bps;
      uint32_t sr;
      std::memcpy(&f16, p, 2); std::memcpy(&ch, p + 2, 2); std::memcpy(&sr, p + 4, 4); std::memcpy(&bps, p + 14, 2);
      fmt = f16; channels = ch; bits = bps; *rate = sr;
      if (fmt == 0xFFFE && len >= 26) { uint16_t sub; std::memcpy(&sub, p + 24, 2); fmt = sub; }   // WAVE_FORMAT_EXTENSIBLE
    } else if (id == "data") {
      const size_t frame = (size_t)channels * bits / 8;
      const size_t n = frame ? len / frame : 0;
      out.resize(n);
      for (size_t i = 0; i < n; ++i) {
        const uint8_t* s = p + i * frame;
        if (fmt == 1 && bits == 16) { int16_t v; std::memcpy(&v, s, 2); out[i] = (float)v / 32768.0f; }
        else if (fmt == 3 && bits == 32) { float v; std::memcpy(&v, s, 4); out[i] = v; }
        else die("unsupported WAV encoding (need PCM16 or float32)");
      }
    }
    pos += 8 + len + (len & 1);
  }
  if (out.empty()) die("WAV file holds no samples: " + path);
  return out;
}

// value of "key": <string|number|[a,b]> in a flat JSON object (the tap sidecar of cpp/include/audio_tap.h:662-721)
bool json_find(const std::string& js, const std::string& key, std::string* val) {
  const size_t k = js.find("\"" + key + "\"");
  if (k == std::string::npos) return false;
  size_t c = js.find(':', k);
  if (c == std::string::npos) return false;
  ++c;
  while (c < js.size() && std::isspace((unsigned char)js[c])) ++c;
  size_t e = c;
  if (js[c] == '"') { e = js.find('"', c + 1); *val = js.substr(c + 1, e - c - 1); return true; }
  if (js[c] == '[') { e = js.find(']', c); *val = js.substr(c + 1, e - c - 1); return true; }
  while (e < js.size() && js[e] != ',' && js[e] != '}' && !std::isspace((unsigned char)js[e])) ++e;
  *val = js.substr(c, e - c);
  return true;
}

std::vector<float> frames_major_to_bins_major(const std::vector<float>& tc, int n_mels, size_t T) {   // main.rs:78-88
  std::vector<float> bct(tc.size());
  for (size_t t = 0; t < T; ++t)
    for (int m = 0; m < n_mels; ++m) bct[(size_t)m * T + t] = tc[t * n_mels + m];
  return bct;
}
std::vector<float> slice_bct(const std::vector<float>& bct, int n_mels, size_t T, size_t start, size_t frames) {   // main.rs:176-185
  std::vector<float> out((size_t)n_mels * frames);
  for (int m = 0; m < n_mels; ++m) std::memcpy(&out[(size_t)m * frames], &bct[(size_t)m * T + start], frames * 4);
  return out;
}
// rust/features/src/lib.rs:127-172 (f32, sequential sums, T-1 denominator, +1e-5 on the std)
void per_feature_stats(const std::vector<float>& tc, size_t T, std::vector<float>* mean, std::vector<float>* stdv) {
  mean->assign(kMels, 0.f); stdv->assign(kMels, 0.f);
  for (int m = 0; m < kMels; ++m) {
    float s = 0.f;
    for (size_t t = 0; t < T; ++t) s += tc[t * kMels + m];
    const float mu = T ? s / (float)T : 0.f;
    float q = 0.f;
    for (size_t t = 0; t < T; ++t) { const float d = tc[t * kMels + m] - mu; q += d * d; }
    (*mean)[m] = mu;
    (*stdv)[m] = std::sqrt(q / (T > 1 ? (float)(T - 1) : 1.0f)) + 1e-5f;
  }
}
void apply_norm(std::vector<float>* tc, size_t T, const std::vector<float>& mean, const std::vector<float>& stdv) {
  for (size_t t = 0; t < T; ++t)
    for (int m = 0; m < kMels; ++m) (*tc)[t * kMels + m] = ((*tc)[t * kMels + m] - mean[m]) / stdv[m];
}

struct Frontend {      // GPU log-mel (replaces rust/features' LogMelExtractor::compute)
  PkbFrontend* f;
  explicit Frontend(int device) : f(pkb_frontend_create(device)) { if (!f) die(std::string("GPU frontend: ") + pkb_last_error()); }
  ~Frontend() { pkb_frontend_destroy(f); }
  std::vector<float> compute(const float* pcm, size_t n, size_t* T) const {
    *T = n < 400 ? 0 : (n - 400) / 160 + 1;
    std::vector<float> out(*T * kMels);
    if (*T && pkb_frontend_logmel(f, pcm, n, out.data(), out.size()) < 0) die(std::string("log-mel: ") + pkb_last_error());
    return out;
  }
};

void drain(ParakeetSession* s, const Args& a, bool streaming_style) {
  ParakeetEvent ev;
  while (parakeet_poll_event(s, &ev)) {
    if (ev.type == PARAKEET_EVENT_FINAL_TEXT) std::printf(streaming_style ? "\nFinal: %s\n" : "Transcript: %s\n", ev.text);
    else if (ev.type == PARAKEET_EVENT_PARTIAL_TEXT) {
      if (a.verbose) std::fprintf(stderr, "[replay] Partial: %s\n", ev.text);
      else if (streaming_style) { std::printf("\rPartial: %s", ev.text); std::fflush(stdout); }
    } else std::fprintf(stderr, "%sError: %s\n", streaming_style ? "\n" : "", ev.error_message);
  }
}

void push(ParakeetSession* s, const std::vector<float>& bct, size_t frames) {
  const int rc = parakeet_push_features(s, bct.data(), frames);
  if (rc < 0) {
    ParakeetEvent ev;
    std::string msg = "push_features failed with error code " + std::to_string(rc);
    for (int i = 0; i < 4 && parakeet_poll_event(s, &ev); ++i)      // rust/parakeet_trt/src/lib.rs:53-70
      if (ev.type == PARAKEET_EVENT_ERROR) { msg += std::string(": ") + ev.error_message; break; }
    die(msg);
  }
}

void dump(const std::vector<float>& v, const std::string& path) {
  std::ofstream f(path, std::ios::binary);
  f.write((const char*)v.data(), (std::streamsize)(v.size() * 4));
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    const std::string s = argv[i];
    auto need = [&](const char* name) -> std::string { if (i + 1 >= argc) die(std::string(name) + " needs a value"); return argv[++i]; };
    if (s == "--model-dir") a.model_dir = need("--model-dir");
    else if (s == "--stream-sim") a.stream_sim = std::atof(need("--stream-sim").c_str());
    else if (s == "--device-id") a.device_id = std::atoi(need("--device-id").c_str());
    else if (s == "--raw-pcm") a.raw_pcm = true;
    else if (s == "--sample-rate") a.sample_rate = std::atol(need("--sample-rate").c_str());
    else if (s == "--features-input") a.features_input = true;
    else if (s == "--n-mels") a.n_mels = std::atoi(need("--n-mels").c_str());
    else if (s == "--verbose" || s == "-v") a.verbose = true;
    else if (s == "--dump-features") a.dump_features = need("--dump-features");
    else if (s == "--feature-norm") a.feature_norm = need("--feature-norm");
    else if (s == "--no-sleep") a.no_sleep = true;      // additive: stream simulation without the real-time sleeps
    else if (s == "--stream-audio") a.stream_audio = std::atof(need("--stream-audio").c_str());
    else if (s == "--whole-utterance") a.whole_utterance = true;   // additive: one full-context pass over the whole file (pkb_offline_utterances)
    else if (s == "--help" || s == "-h") {
      std::printf("usage: parakeet_cli <input> --model-dir DIR [--stream-sim SEC] [--device-id N] [--raw-pcm] [--sample-rate HZ]\n"
                  "       [--features-input] [--n-mels N] [-v|--verbose] [--dump-features PATH] [--feature-norm none|per_feature] [--no-sleep]\n"
                  "       [--whole-utterance] [--stream-audio SECONDS]\n"
                  "  --stream-sim S     reference behaviour (rust/cli): every S seconds of audio become ONE independent push of its own frames\n"
                  "  --stream-audio S   cache-aware streaming: audio is pushed S seconds at a time, the engine frames it and cuts the 41 / 57-frame\n"
                  "                     chunk schedule itself (nothing is dropped between pushes); --feature-norm running = causal running statistics\n");
      return 0;
    } else if (!s.empty() && s[0] == '-') die("unknown option " + s);
    else a.input = s;
  }
  if (a.input.empty() || a.model_dir.empty()) die("usage: parakeet_cli <input> --model-dir DIR [options]   (--help)");
  std::string norm = a.feature_norm;
  if (norm.empty()) { const char* e = std::getenv("PARAKEET_FEATURE_NORM"); norm = e ? e : "none"; }
  if (norm != "none" && norm != "per_feature" && !(norm == "running" && a.stream_audio > 0)) die("Unsupported feature normalization: " + norm);
  const bool per_feature = norm == "per_feature";
  const auto t_start = std::chrono::steady_clock::now();
  if (a.verbose) {
    std::fprintf(stderr, "[replay] Input: %s\n[replay] Model: %s\n[replay] Mode: %s\n[replay] Feature normalization: %s\n", a.input.c_str(),
                 a.model_dir.c_str(), a.features_input ? "features" : a.raw_pcm ? "raw_pcm" : "wav", norm.c_str());
  }
  ParakeetConfig cfg{a.model_dir.c_str(), a.device_id, true};

  if (a.features_input) {      // main.rs:208-334
    std::string raw = a.input, json;
    const bool is_json = raw.size() > 5 && raw.substr(raw.size() - 5) == ".json";
    if (is_json) { json = raw; raw = raw.substr(0, raw.size() - 5) + ".raw"; }
    else { const size_t dot = raw.rfind('.'); json = (dot == std::string::npos ? raw : raw.substr(0, dot)) + ".json"; }
    int n_mels = a.n_mels > 0 ? a.n_mels : kMels;
    std::string layout = "bins_major";
    std::ifstream jf(json);
    if (jf) {
      const std::string js((std::istreambuf_iterator<char>(jf)), std::istreambuf_iterator<char>());
      std::string v;
      if (json_find(js, "format", &v) && v != "f32le") die("Feature JSON format '" + v + "' not supported (expected f32le)");
      if (json_find(js, "layout", &v)) { layout.clear(); for (char c : v) if (!std::isspace((unsigned char)c)) layout.push_back((char)std::tolower(c)); }
      if (a.n_mels <= 0) {
        if (json_find(js, "mel_bins", &v)) n_mels = std::atoi(v.c_str());
        else if (json_find(js, "shape", &v)) {
          const int d0 = std::atoi(v.c_str()), d1 = std::atoi(v.substr(v.find(',') + 1).c_str());
          n_mels = layout == "frames_major" ? d1 : d0;
        }
      }
    } else if (is_json) die("Feature JSON not found: " + json);
    if (n_mels != kMels) die("this model takes 128 mel bins (got " + std::to_string(n_mels) + ")");
    std::vector<float> feats = load_f32le(raw, "Feature");
    if (feats.size() % n_mels != 0) die("Feature file size not divisible by n_mels");
    const size_t T = feats.size() / n_mels;
    if (layout == "frames_major") feats = frames_major_to_bins_major(feats, n_mels, T);
    if (a.verbose) std::fprintf(stderr, "[replay] Loaded %zu frames of %d mel features\n", T, n_mels);
    ParakeetSession* s = parakeet_create_session(&cfg);
    if (!s) die("session creation failed");
    std::printf("Starting transcription (feature replay)...\n");
    for (size_t start = 0, idx = 0; start < T; start += 256, ++idx) {
      const size_t frames = std::min<size_t>(256, T - start);
      if (a.verbose && T > 256) std::fprintf(stderr, "[replay] feature_chunk=%zu start=%zu frames=%zu\n", idx, start, frames);
      push(s, T <= 256 ? feats : slice_bct(feats, n_mels, T, start, frames), frames);
      drain(s, a, false);
    }
    parakeet_destroy_session(s);
  } else {
    long rate = a.sample_rate > 0 ? a.sample_rate : 16000;
    std::vector<float> audio = a.raw_pcm ? load_f32le(a.input, "Raw PCM") : load_wav(a.input, &rate);
    if (rate != 16000) die("expected 16 kHz audio (got " + std::to_string(rate) + " Hz); resample first");
    if (a.verbose) std::fprintf(stderr, "[replay] Loaded %zu samples (%.2f s)\n", audio.size(), audio.size() / 16000.0);
    if (a.whole_utterance) {
      // The offline file mode of the reference CLI pushes the whole file at once (main.rs:484-535) and the runtime cuts it into
      // independent <= 256-frame segments; this mode keeps the utterance whole: log-mel, normalisation, self-attention over ALL
      // frames and the TDT loop in one pkb_offline_utterances call.
      if (audio.size() < 400) die("audio shorter than one 25 ms frame");
      const int frames = (int)((audio.size() - 400) / 160 + 1);
      PkbEngineConfig ec{};
      ec.model_dir = a.model_dir.c_str(); ec.device_id = a.device_id; ec.max_streams = 1; ec.precision = 0; ec.gemm_backend = 0;
      ec.contract_cache = 0; ec.punct_suppression = 1; ec.max_rows = pkb_encoded_length(frames) + 64;
      PkbEngine* eng = pkb_engine_create(&ec);
      if (!eng) die(std::string("engine creation failed: ") + pkb_last_error());
      const int32_t sid = pkb_stream_open(eng);
      const float* ap = audio.data();
      const size_t ns = audio.size();
      std::printf("Starting transcription (whole utterance)...\n");
      if (sid < 0 || pkb_offline_utterances(eng, 1, &sid, &ap, &ns, per_feature ? 1 : 0, nullptr, nullptr, 0, nullptr, 1) != 0)
        die(std::string("pkb_offline_utterances failed: ") + pkb_last_error());
      std::vector<char> text((size_t)pkb_stream_text(eng, sid, nullptr, 0) + 1);
      pkb_stream_text(eng, sid, text.data(), (int32_t)text.size());
      if (a.verbose) std::fprintf(stderr, "[replay] %d feature frames -> %d encoder frames, %d tokens\n", frames, pkb_encoded_length(frames),
                                  pkb_stream_num_tokens(eng, sid));
      std::printf("Transcript: %s\n", text.data());
      pkb_engine_destroy(eng);
      if (a.verbose)
        std::fprintf(stderr, "[replay] Completed in %.2fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
      return 0;
    }
    if (a.stream_audio > 0) {
      // Additive mode: the reference CLI's --stream-sim makes every interval an independent push, of which the streaming
      // encoder decodes only valid_out_len = 3 frames -- fine for replaying 48-frame chunks, lossy for anything else.  Here the audio goes
      // through pkb_stream_push_audio: the engine keeps the sample / frame carry-over, cuts the cache-aware schedule (41 frames, then 57-frame
      // slices shifted by 24) and decodes every encoder frame once.  The tail is flushed with up to one chunk of zero samples.
      const size_t per_push = (size_t)(a.stream_audio * 16000.0);
      if (per_push == 0) die("--stream-audio interval too small");
      PkbEngineConfig ec{};
      ec.model_dir = a.model_dir.c_str(); ec.device_id = a.device_id; ec.max_streams = 1; ec.precision = 0; ec.gemm_backend = 0;
      ec.contract_cache = 0; ec.punct_suppression = 1; ec.max_rows = 0;
      PkbEngine* eng = pkb_engine_create(&ec);
      if (!eng) die(std::string("engine creation failed: ") + pkb_last_error());
      const int32_t sid = pkb_stream_open(eng);
      if (sid < 0) die(std::string("pkb_stream_open failed: ") + pkb_last_error());
      if (norm == "running") {
        if (pkb_stream_set_feature_norm_running(eng, sid, 1) != 0) die(std::string("running normalisation: ") + pkb_last_error());
      } else if (per_feature) {      // whole-file statistics, like the reference's stream-sim (main.rs:398-405)
        Frontend fe(a.device_id);
        size_t T = 0;
        std::vector<float> tc = fe.compute(audio.data(), audio.size(), &T);
        std::vector<float> mean, stdv;
        per_feature_stats(tc, T, &mean, &stdv);
        if (pkb_stream_set_feature_norm(eng, sid, mean.data(), stdv.data()) != 0) die(std::string("feature norm: ") + pkb_last_error());
      }
      std::printf("Starting transcription (cache-aware streaming from audio)...\n");
      std::vector<char> text;
      auto drain_steps = [&](bool final_print) {
        while (pkb_stream_has_pending(eng, sid) > 0) {
          const int n = pkb_engine_step(eng);
          if (n < 0) die(std::string("pkb_engine_step failed: ") + pkb_last_error());
          if (n == 0) break;
        }
        text.assign((size_t)pkb_stream_text(eng, sid, nullptr, 0) + 1, 0);
        pkb_stream_text(eng, sid, text.data(), (int32_t)text.size());
        if (final_print) std::printf("\nFinal: %s\n", text.data());
        else { std::printf("\rPartial: %s", text.data()); std::fflush(stdout); }
      };
      for (size_t pos = 0; pos < audio.size(); pos += per_push) {
        const size_t n = std::min(per_push, audio.size() - pos);
        if (pkb_stream_push_audio(eng, sid, audio.data() + pos, n) != 0) die(std::string("pkb_stream_push_audio failed: ") + pkb_last_error());
        drain_steps(false);
        if (!a.no_sleep) std::this_thread::sleep_for(std::chrono::duration<double>(a.stream_audio));
      }
      const std::vector<float> zeros((size_t)57 * 160 + 400, 0.0f);      // completes the last partial chunk of the schedule
      if (pkb_stream_push_audio(eng, sid, zeros.data(), zeros.size()) != 0) die(std::string("flush failed: ") + pkb_last_error());
      drain_steps(true);
      if (a.verbose)
        std::fprintf(stderr, "[replay] %lld chunks, %lld encoder frames, %d tokens\n", (long long)pkb_stream_chunks_done(eng, sid),
                     (long long)pkb_stream_encoder_frames(eng, sid), pkb_stream_num_tokens(eng, sid));
      pkb_engine_destroy(eng);
      if (a.verbose)
        std::fprintf(stderr, "[replay] Completed in %.2fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
      return 0;
    }
    Frontend fe(a.device_id);
    std::vector<float> mean, stdv;
    std::vector<float> whole_tc;
    size_t whole_T = 0;
    if (per_feature || a.stream_sim < 0) whole_tc = fe.compute(audio.data(), audio.size(), &whole_T);
    if (per_feature) per_feature_stats(whole_tc, whole_T, &mean, &stdv);      // stats over the whole file even in stream-sim (main.rs:398-405)
    ParakeetSession* s = parakeet_create_session(&cfg);
    if (!s) die("session creation failed");
    std::printf("Starting transcription...\n");
    std::vector<float> all_bct;
    if (a.stream_sim > 0) {      // main.rs:416-474
      const size_t per_chunk = (size_t)(a.stream_sim * 16000.0);
      if (per_chunk == 0) die("--stream-sim interval too small");
      for (size_t pos = 0, idx = 0; pos < audio.size(); pos += per_chunk, ++idx) {
        const size_t n = std::min(per_chunk, audio.size() - pos);
        size_t T = 0;
        std::vector<float> tc = fe.compute(audio.data() + pos, n, &T);
        if (T > 0 && T < 33) {
          // the streaming encoder needs 33 frames per push (8x subsampling, 2 dropped + 3 emitted tokens): a shorter tail is skipped
          std::fprintf(stderr, "[replay] WARNING: chunk=%zu has %zu frames (< 33, the streaming encoder's minimum): skipped\n", idx, T);
        } else if (T > 0) {
          if (per_feature) apply_norm(&tc, T, mean, stdv);
          const std::vector<float> bct = frames_major_to_bins_major(tc, kMels, T);
          if (!a.dump_features.empty()) all_bct.insert(all_bct.end(), bct.begin(), bct.end());
          if (a.verbose) std::fprintf(stderr, "[replay] chunk=%zu pos=%zu samples=%zu frames=%zu\n", idx, pos, n, T);
          push(s, bct, T);
          drain(s, a, true);
        }
        if (!a.no_sleep) std::this_thread::sleep_for(std::chrono::duration<double>(a.stream_sim));
      }
      std::printf("\n");
      if (!a.dump_features.empty()) dump(all_bct, a.dump_features);
    } else {      // offline whole file, one push (main.rs:484-535)
      if (per_feature) apply_norm(&whole_tc, whole_T, mean, stdv);
      const std::vector<float> bct = frames_major_to_bins_major(whole_tc, kMels, whole_T);
      if (a.verbose) std::fprintf(stderr, "[replay] Computed %zu feature frames\n", whole_T);
      if (!a.dump_features.empty()) dump(bct, a.dump_features);
      push(s, bct, whole_T);
      drain(s, a, false);
    }
    parakeet_destroy_session(s);
  }
  if (a.verbose)
    std::fprintf(stderr, "[replay] Completed in %.2fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
  return 0;
}
