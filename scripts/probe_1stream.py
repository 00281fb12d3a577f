"""One stream, steady state: a handful of chunk steps for an ncu launch list (per-kernel durations at the latency configuration)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "trt-asr-engine_b200"), os.path.join(ROOT, "trt-asr-engine_b200", "tools")):
    sys.path.insert(0, p)
import numpy as np
import binding
from make_synthetic_model import ensure_model
from synth_audio import synth_clip
_L = int(os.environ.get("LAYERS", "24"))
model = ensure_model(os.path.join(ROOT, "models", f"synth{_L}"), n_layers=_L, seed=0)
n = int(os.environ.get("STREAMS", "1"))
eng = binding.Engine(model, max_streams=n, precision=int(os.environ.get("PREC", "0")))
sids = np.array([eng.open() for _ in range(n)], np.int32)
clip = synth_clip(8.0, 1000)
S = 3840
buf = np.ascontiguousarray(np.stack([clip[:16 * S]] * n))
for k in range(int(os.environ.get("CHUNKS", "8"))):
    seg = np.ascontiguousarray(buf[:, (k % 16) * S:(k % 16 + 1) * S])
    eng.push_audio_batch(sids, seg.ctypes.data, S, S)
    eng.step()
print("launches", eng.kernel_launches(), "tokens", eng.tokens(int(sids[0])))
