"""Runs the legacy session (offline encoder mode, fp32-grade arithmetic, reference-format decode trace on stderr) over the clip of
tests/golden/tdt_trace_ref.json, so that the reference's own tools/verify_nemo/compare_tdt_trace.py can be run (in the build
container) on this library's trace against the reference's PyTorch-loop trace.  stderr -> gpurun_out/tdt_steps_b200_stderr.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "trt-asr-engine_b200"), os.path.join(ROOT, "trt-asr-engine_b200", "tools"),
          os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
os.environ.update(PARAKEET_DEBUG_TDT_STEPS="100000", PARAKEET_B200_ENCODER="offline", PARAKEET_DISABLE_PUNCT_SUPPRESSION="1")
import json  # noqa: E402

import numpy as np  # noqa: E402

import binding  # noqa: E402
from conftest import FeaturesRef, build_oracle, model_dir, normalized_features  # noqa: E402

doc = json.load(open(os.path.join(ROOT, "tests", "golden", "tdt_trace_ref.json")))
f = normalized_features(FeaturesRef(build_oracle()), doc["clip"]["seconds"], doc["clip"]["seed"])
f[0] = 0.0
s = binding.ParakeetSessionSafe(model_dir(2), 0, use_fp16=False)
C = doc["chunk_frames"]
for lo in range(0, f.shape[1], C):
    seg = np.ascontiguousarray(f[:, lo:lo + C])
    s.push_features(seg, seg.shape[1])
s.close()
