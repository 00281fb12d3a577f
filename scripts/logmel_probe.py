"""Log-mel frontend on one 1 h clip (BASELINE config 5's frontend input): the launch an ncu capture of the frontend kernel targets,
and a direct A/B of the two implementations (PARAKEET_B200_LOGMEL=0: shared-memory Stockham FFT)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "trt-asr-engine_b200"), os.path.join(ROOT, "trt-asr-engine_b200", "tools")):
    sys.path.insert(0, p)
import numpy as np
import binding
from make_synthetic_model import ensure_model
from synth_audio import synth_clip
model = ensure_model(os.path.join(ROOT, "models", "synth2"), n_layers=2, seed=0)
eng = binding.Engine(model, max_streams=1, precision=0)
hour = np.tile(synth_clip(10.0, 1234), 360)
eng.logmel(hour[:160000])
eng.profile_enable(True)
for _ in range(int(os.environ.get("REPS", "3"))):
    out = eng.logmel(hour)
ms, nbytes, n = eng.profile_read_class(2)
print(f"logmel 1 h clip: {out.shape[0]} frames, {ms / n:.3f} ms per launch, {nbytes / ms / 1e6:.0f} GB/s algorithmic, impl env={os.environ.get('PARAKEET_B200_LOGMEL', 'default')}")
eng.close()
