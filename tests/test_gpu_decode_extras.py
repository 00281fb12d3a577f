"""The token-changing extras the reference applies INSIDE its greedy loop, on both decode paths of the product (<= 16 streams:
CUDA-core joint + full-logit scan in tdt_select_kernel; > 16 streams: tcgen05 joint with the argmax fused into its epilogue):
  * PARAKEET_BLANK_PENALTY subtracted from the blank logit      (/root/reference/cpp/src/parakeet_trt.cpp:3175-3178)
  * NaN logits read as -100 on both heads                        (:2971)
  * leading punctuation-only piece -> blank while nothing has been emitted in the utterance (:3256-3262), and its off switch
    PARAKEET_DISABLE_PUNCT_SUPPRESSION (:2872)
Each case first shows on the oracle that the extra CHANGES the trace of the clip (the test is not vacuous), then requires the GPU
trace to equal the oracle's, chunk by chunk, closed loop, fp32-grade arithmetic (precision 1)."""
import os
import shutil
import struct

import numpy as np
import pytest
import torch

import binding
from conftest import normalized_features
from model_ref import DecodeState, ModelRef, prime, streaming_schedule, tdt_greedy_chunk
from weights_io import _CFG, _HDR, _TEN

pytestmark = pytest.mark.gpu
N_CHUNKS = 14
PATHS = [pytest.param(1, id="scan-path-1-stream"), pytest.param(20, id="fused-argmax-20-streams")]


def _feats(features_ref):
    f = normalized_features(features_ref, 0.41 + 0.24 * N_CHUNKS + 0.5, 4242)
    f[0] = 0.0
    return f


def _oracle_traces(m, f, **kw):
    st = DecodeState(m)
    prime(m, st)
    cc, ct, cl = m.initial_cache(1)
    out = []
    for b, e in streaming_schedule(N_CHUNKS):
        enc, el, cc, ct, cl = m.stream_step(torch.from_numpy(f[None, :, b:e]), torch.tensor([e - b]), cc, ct, cl)
        out.append([(t, tok, d) for t, tok, d, _ in tdt_greedy_chunk(m, st, enc, int(el), **kw)])
    return out, st.tokens


def _gpu_traces(model, f, n_streams, **kw):
    eng = binding.Engine(model, max_streams=n_streams, precision=1, **kw)
    sids = [eng.open() for _ in range(n_streams)]
    out = [[] for _ in sids]
    for b, e in streaming_schedule(N_CHUNKS):
        for s in sids:
            eng.push_features(s, f[:, b:e])
        assert eng.step() == n_streams
        for i, s in enumerate(sids):
            out[i].append(eng.last_steps(s))
    eng.close()
    assert all(o == out[0] for o in out), "copies of one clip in one batch disagree"
    return out[0]


def _tensor_offset(path, name):
    raw = open(path, "rb").read(1 << 20)
    _, _, n_t, n_c, _ = _HDR.unpack_from(raw, 0)
    pos = _HDR.size + _CFG.size * n_c
    for _ in range(n_t):
        nm, dt, nd, d0, d1, d2, d3, off, nb = _TEN.unpack_from(raw, pos)
        pos += _TEN.size
        if nm.rstrip(b"\0").decode() == name:
            return off, nb, dt
    raise KeyError(name)


def _variant(tmp_path, src, vocab=None, patch_bias=None):
    d = tmp_path / "variant"
    d.mkdir()
    if patch_bias:
        shutil.copyfile(os.path.join(src, "weights.bin"), d / "weights.bin")
        off, nb, dt = _tensor_offset(str(d / "weights.bin"), "joint.joint_net.2.bias")
        assert dt == 0 and nb == 8198 * 4
        with open(d / "weights.bin", "r+b") as fh:
            for idx, val in patch_bias.items():
                fh.seek(off + 4 * idx)
                fh.write(struct.pack("<f", val))
    else:
        os.symlink(os.path.join(src, "weights.bin"), d / "weights.bin")
    lines = vocab if vocab is not None else open(os.path.join(src, "vocab.txt"), encoding="utf-8").read().split("\n")[:-1]
    with open(d / "vocab.txt", "w", encoding="utf-8") as fh:
        fh.write("\n".join(lines) + "\n")
    return str(d)


@pytest.mark.parametrize("n_streams", PATHS)
@pytest.mark.parametrize("penalty", [2.5, -1.5])
def test_blank_penalty(model_small, oracle_small, features_ref, n_streams, penalty):
    f = _feats(features_ref)
    base, _ = _oracle_traces(oracle_small, f)
    want, _ = _oracle_traces(oracle_small, f, blank_penalty=penalty)
    assert want != base, "penalty does not change this clip's trace: pick another clip"
    os.environ["PARAKEET_BLANK_PENALTY"] = str(penalty)
    try:
        got = _gpu_traces(model_small, f, n_streams)
    finally:
        del os.environ["PARAKEET_BLANK_PENALTY"]
    assert got == want


@pytest.mark.parametrize("n_streams", PATHS)
def test_nan_logits_read_as_minus_100(tmp_path, model_small, oracle_small, features_ref, n_streams):
    f = _feats(features_ref)
    base, toks = _oracle_traces(oracle_small, f)
    assert toks, "clip emits nothing"
    top_tok = max(set(toks), key=toks.count)
    durs = [d for ch in base for _, _, d in ch]
    top_dur = max(set(durs), key=durs.count)
    model = _variant(tmp_path, model_small, patch_bias={top_tok: float("nan"), 8193 + top_dur: float("nan")})
    m = ModelRef(model)
    assert np.isnan(m.w["joint.joint_net.2.bias"][top_tok].item())
    want, toks2 = _oracle_traces(m, f)
    assert want != base and top_tok not in toks2 and all(d != top_dur for ch in want for _, _, d in ch)
    assert _gpu_traces(model, f, n_streams) == want


@pytest.mark.parametrize("n_streams", PATHS)
def test_leading_punctuation_suppression(tmp_path, model_small, oracle_small, features_ref, n_streams):
    f = _feats(features_ref)
    base, toks = _oracle_traces(oracle_small, f, punct_suppression=False)
    assert len(toks) >= 2
    vocab = list(oracle_small.vocab_lines)
    assert not oracle_small.is_punct_only(toks[0])
    vocab[toks[0]] = "▁,"                      # the clip's first emission becomes a punctuation-only piece
    model = _variant(tmp_path, model_small, vocab=vocab)
    m = ModelRef(model)
    assert m.is_punct_only(toks[0])
    want, toks_s = _oracle_traces(m, f)            # suppression on (the default)
    assert want != base and toks_s and toks_s[0] != toks[0]
    assert _gpu_traces(model, f, n_streams) == want
    # PARAKEET_DISABLE_PUNCT_SUPPRESSION: the same model decodes like the unmodified vocabulary
    assert _gpu_traces(model, f, n_streams, punct_suppression=0) == base
