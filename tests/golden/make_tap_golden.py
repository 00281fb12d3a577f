#!/usr/bin/env python3
"""Generates tests/golden/tap_features_{bins,frames}_major.{raw,json}: feature taps written by the REFERENCE's own
FeatureTapWriter (/root/reference/cpp/include/audio_tap.h, compiled by oracle/Makefile into oracle/_ref/feature_tap_writer).
Run in the build container (the reference tree is not on the GPU box); the fixtures (48 frames x 128 mel bins) are committed."""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "trt-asr-engine_b200"), os.path.join(ROOT, "oracle")]
from conftest import FeaturesRef, build_oracle  # noqa: E402
from synth_audio import synth_clip  # noqa: E402

fr = FeaturesRef(build_oracle())
feat = fr.normalized(fr.logmel(synth_clip(0.5, 99)))[:48]          # [48, 128] frames-major
exe = os.path.join(ROOT, "oracle", "_ref", "feature_tap_writer")
for layout, data in (("bins_major", np.ascontiguousarray(feat.T)), ("frames_major", np.ascontiguousarray(feat))):
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "in.f32")
        data.astype(np.float32).tofile(src)
        env = dict(os.environ, AUDIO_TAP_ENABLE="1", AUDIO_TAP_DIR=d, AUDIO_TAP_FEATURES="1")
        run_dir = subprocess.run([exe, src, str(feat.shape[0]), layout], env=env, check=True, capture_output=True, text=True).stdout.strip()
        (raw,) = glob.glob(os.path.join(run_dir, "tap_FEATURES.raw"))
        for ext in ("raw", "json"):
            shutil.copy(raw[:-3] + ext, os.path.join(ROOT, "tests", "golden", f"tap_features_{layout}.{ext}"))
print(open(os.path.join(ROOT, "tests", "golden", "tap_features_bins_major.json")).read())
