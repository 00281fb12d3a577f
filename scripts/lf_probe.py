"""Small driver for profiling the whole-utterance attention kernel: one synthetic clip through pkb_offline_utterances."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "trt-asr-engine_b200"), os.path.join(ROOT, "trt-asr-engine_b200", "tools")):
    sys.path.insert(0, p)
import binding
from make_synthetic_model import ensure_model
from synth_audio import synth_clip

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 1200.0
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = ensure_model(os.path.join(ROOT, "models", f"synth{layers}"), n_layers=layers, seed=0)
n = int(seconds * 16000)
pcm = np.tile(synth_clip(10.0, 1234), n // 160000 + 1)[:n]
t_enc = binding.load_library().pkb_encoded_length((n - 400) // 160 + 1)
eng = binding.Engine(model, max_streams=1, precision=0, max_rows=t_enc + 64, contract_cache=0)
s = eng.open()
for rep in range(2):
    eng.profile_enable(True)
    t0 = time.perf_counter()
    eng.offline_utterances([s], audio=[pcm], decode=False)
    wall = time.perf_counter() - t0
    ms, fl, l = eng.profile_read_class(4)
    eng.profile_enable(False)
    print(f"rep {rep}: T_enc={t_enc} wall={wall*1e3:.1f} ms attention {ms/max(l,1):.3f} ms/launch {fl/max(ms,1e-9)/1e9:.1f} TFLOP/s (algorithmic)")
    eng.reset(s)
