"""Small driver for checking / profiling the persistent decode loop: one or more synthetic clips through pkb_offline_utterances with decode on."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "trt-asr-engine_b200"), os.path.join(ROOT, "trt-asr-engine_b200", "tools")):
    sys.path.insert(0, p)
import binding
from make_synthetic_model import ensure_model
from synth_audio import synth_clip

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nclips = int(sys.argv[3]) if len(sys.argv) > 3 else 2
model = ensure_model(os.path.join(ROOT, "models", f"synth{layers}"), n_layers=layers, seed=0)
n = int(seconds * 16000)
clips = [np.tile(synth_clip(10.0, 1234 + i), n // 160000 + 1)[:n] for i in range(nclips)]
t_enc = binding.load_library().pkb_encoded_length((n - 400) // 160 + 1)
eng = binding.Engine(model, max_streams=nclips, precision=0, max_rows=nclips * (t_enc + 64), contract_cache=0)
sids = [eng.open() for _ in range(nclips)]
t0 = time.perf_counter()
eng.offline_utterances(sids, audio=clips)
print(f"T_enc={t_enc} clips={nclips} wall={1e3*(time.perf_counter()-t0):.1f} ms tokens={[len(eng.tokens(s)) for s in sids]}")
