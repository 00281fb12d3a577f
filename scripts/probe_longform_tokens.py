"""How many tokens does the random-weight model emit on long clips as a function of PARAKEET_BLANK_PENALTY (bench config 5 calibration)."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "trt-asr-engine_b200"), os.path.join(ROOT, "trt-asr-engine_b200", "tools")):
    sys.path.insert(0, p)
import numpy as np
import binding
from make_synthetic_model import ensure_model
from synth_audio import synth_clip
model = ensure_model(os.path.join(ROOT, "models", "synth24"), n_layers=24, seed=0)
segs = [synth_clip(10.0, 4000 + k) for k in range(12)]
rng = np.random.default_rng(7)
for secs in (10, 60, 600, 3600):
    n_samp = secs * 16000
    order = rng.integers(0, len(segs), size=n_samp // 160000 + 1)
    audio = np.concatenate([segs[k] for k in order])[:n_samp].astype(np.float32)
    t_enc = binding.load_library().pkb_encoded_length((n_samp - 400) // 160 + 1)
    eng = binding.Engine(model, max_streams=1, precision=0, max_rows=t_enc + 64, contract_cache=0)
    s = eng.open()
    for pen in (0.0, 4.0, 8.0, 12.0, 16.0, 24.0):
        eng.set_blank_penalty(pen)
        eng.reset(s)
        t0 = time.perf_counter()
        eng.offline_utterances([s], audio=[audio], per_feature_norm=True, decode=True)
        dt = time.perf_counter() - t0
        print(f"secs={secs} penalty={pen} tokens={len(eng.tokens(s))} steps={len(eng.last_steps(s))} wall={dt:.3f}", flush=True)
    eng.close()
