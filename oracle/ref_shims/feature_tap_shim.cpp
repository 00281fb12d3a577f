// feature_tap_shim.cpp -- TEST INFRASTRUCTURE.  Drives the REFERENCE's own FeatureTapWriter
// (/root/reference/cpp/include/audio_tap.h:600-780, header-only) so that tests/golden holds a feature tap (raw + JSON sidecar)
// exactly as the reference runtime writes it; the CLI's sidecar parser (rust/cli/src/main.rs:132-165 semantics) is tested on it.
//   usage: AUDIO_TAP_ENABLE=1 AUDIO_TAP_DIR=<dir> AUDIO_TAP_FEATURES=1 feature_tap_writer <in.f32> <n_frames> <bins_major|frames_major>
// The input holds n_frames x 128 floats already in the requested layout; it is written in two write_frames() calls.
#include <limits>
#include <cstdio>
#include <vector>

#include "audio_tap.h"

int main(int argc, char** argv) {
  if (argc != 4) { std::fprintf(stderr, "usage: %s in.f32 n_frames layout\n", argv[0]); return 2; }
  const size_t T = (size_t)std::atol(argv[2]);
  const std::string layout = argv[3];
  std::vector<float> x(T * 128);
  FILE* f = std::fopen(argv[1], "rb");
  if (!f || std::fread(x.data(), 4, x.size(), f) != x.size()) { std::fprintf(stderr, "short read\n"); return 1; }
  std::fclose(f);
  {
    audio_tap::FeatureTapWriter tap("features", 128, 10.0f, 25.0f, 16000, layout, "golden fixture written by the reference's FeatureTapWriter");
    if (!tap.is_enabled()) { std::fprintf(stderr, "tap not enabled (AUDIO_TAP_ENABLE / AUDIO_TAP_FEATURES)\n"); return 1; }
    tap.write_frames(x.data(), T);      // one call: a bins-major tap is [128, T] of the whole call
  }
  std::printf("%s\n", audio_tap::TapConfig::instance().run_dir.c_str());
  return 0;
}
