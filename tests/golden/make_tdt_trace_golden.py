#!/usr/bin/env python3
"""Generates tests/golden/tdt_trace_ref.json by EXECUTING the reference's own TDT decode loop, unmodified:
/root/reference/tools/verify_nemo/tdt_trace.py (main(): priming with <|startoftranscript|> / <|en|>, per-chunk offline encoder,
per-frame symbol loop, argmax over the two heads, blank+duration-0 clamp, duration advance, max_symbols, leftover advance dropped
at the chunk end).  The script imports NeMo only to obtain a model object; here a stand-in `nemo.collections.asr` hands it the
oracle's modules (oracle/model_ref.py: offline encoder, predictor, joint) for the seeded 2-layer synthetic model, so every
control-flow decision in the resulting trace is the REFERENCE's.  tests/test_oracle_kats.py then requires the oracle's own loop
(tdt_greedy_chunk) to reproduce that trace.  Run in the build container (the reference tree is not on the GPU box)."""
import json
import os
import runpy
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "trt-asr-engine_b200"), os.path.join(ROOT, "trt-asr-engine_b200", "tools"),
                os.path.join(ROOT, "oracle")]
from conftest import FeaturesRef, build_oracle, model_dir  # noqa: E402
from model_ref import ModelRef  # noqa: E402
from synth_audio import synth_clip  # noqa: E402

REF = "/root/reference"
MODEL_DIR = model_dir(2)
SECONDS, SEED, CHUNK = 10.0, 1234, 256


class _Cfg:
    class preprocessor:
        @staticmethod
        def get(key, default=None):
            return 128 if key == "features" else default


class _Decoder:
    def __init__(self, m):
        self.m = m

    def predict(self, y, state, add_sos=False):
        g, h, c = self.m.predictor_step(y, state[0], state[1])
        return g.transpose(1, 2), [h, c]                  # NeMo returns g as [B, U, H]; the script transposes it back


class _Joint(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, encoder_outputs, decoder_outputs):
        return self.m.joint_logits(encoder_outputs, decoder_outputs)


class _Model:
    def __init__(self, m):
        self.m, self.cfg, self.decoder, self.joint = m, _Cfg, _Decoder(m), _Joint(m)

    def eval(self):
        return self

    def to(self, device):
        return self

    def encoder(self, x=None, x_len=None, audio_signal=None, length=None):
        return self.m.offline(x if x is not None else audio_signal, x_len if x_len is not None else length)


def main():
    m = ModelRef(MODEL_DIR)
    nemo = types.ModuleType("nemo")
    coll = types.ModuleType("nemo.collections")
    asr = types.ModuleType("nemo.collections.asr")
    asr.models = types.SimpleNamespace(ASRModel=types.SimpleNamespace(restore_from=lambda p: _Model(m), from_pretrained=lambda p: _Model(m)),
                                       EncDecRNNTBPEModel=types.SimpleNamespace(restore_from=lambda p: _Model(m)))
    nemo.collections, coll.asr = coll, asr
    sys.modules.update({"nemo": nemo, "nemo.collections": coll, "nemo.collections.asr": asr})
    fr = FeaturesRef(build_oracle())
    feat = fr.normalized(fr.logmel(synth_clip(SECONDS, SEED)))            # [T, 128] time-major, the layout the script loads
    feat[:, 0] = 0.0
    with tempfile.TemporaryDirectory() as d:
        fpath, out = os.path.join(d, "f.f32"), os.path.join(d, "trace.jsonl")
        feat.astype(np.float32).tofile(fpath)
        sys.argv = ["tdt_trace.py", "--model", MODEL_DIR, "--model-dir", MODEL_DIR, "--features-f32", fpath, "--chunk-frames", str(CHUNK),
                    "--contract", os.path.join(REF, "contracts", "parakeet-tdt-0.6b-v3.contract.json"), "--out", out]
        try:
            runpy.run_path(os.path.join(REF, "tools", "verify_nemo", "tdt_trace.py"), run_name="__main__")
        except SystemExit as e:
            assert not e.code, e.code
        lines = [json.loads(ln) for ln in open(out)]
        if os.environ.get("KEEP_JSONL"):      # the raw reference-format trace, e.g. for tools/verify_nemo/compare_tdt_trace.py
            import shutil
            shutil.copy(out, os.environ["KEEP_JSONL"])
    meta = lines[0]
    steps = [{"chunk_idx": r["chunk_idx"], "time_idx": r["time_idx"], "u": r["u"], "y_id": r["y_id"], "best_tok": r["best_tok"],
              "best_dur_idx": r["best_dur_idx"], "duration": r["duration"], "advance": r["advance"],
              "tok_gap": r["tok_topk"][0]["v"] - r["tok_topk"][1]["v"], "dur_gap": r["dur_topk"][0]["v"] - r["dur_topk"][1]["v"]}
             for r in lines[1:]]
    doc = {"source": "reference tools/verify_nemo/tdt_trace.py executed on the oracle's modules (tests/golden/make_tdt_trace_golden.py)",
           "model": "synth2 (seed 0)", "clip": {"seconds": SECONDS, "seed": SEED, "mel0_zeroed": True}, "chunk_frames": CHUNK, "meta": meta, "steps": steps}
    json.dump(doc, open(os.path.join(ROOT, "tests", "golden", "tdt_trace_ref.json"), "w"), indent=0)
    print("steps", len(steps), "emitted", sum(1 for s in steps if s["best_tok"] != meta["blank_id"]), "meta", meta)


if __name__ == "__main__":
    main()
