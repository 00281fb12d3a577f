"""Whole-utterance offline path (pkb_offline_utterances; BASELINE configs 1 and 5) on the GPU vs the oracle's offline():
full self-attention over ALL frames of an utterance, symmetric conv padding, greedy TDT over every encoder frame.

Tolerances: those of test_gpu_model.py (precise: 2e-3; bf16: p95 3e-2 / max 1.5e-1, x2 as for the <=256-frame offline engine).
Decode traces: precise mode identical; bf16 mode identical up to the first decision whose top-2 logit gap is below TAU (a
decision that close is ambiguous for any bf16 implementation, see test_gpu_parity_set.py), and that prefix must be non-trivial.
"""
import numpy as np
import pytest
import torch

import binding
from conftest import normalized_features
from model_ref import DecodeState, ModelRef, prime, tdt_greedy_chunk
from synth_audio import synth_clip

pytestmark = pytest.mark.gpu
TAU = 0.25


def _feats(features_ref, seconds, seed):
    f = normalized_features(features_ref, seconds, seed)
    f[0] = 0.0    # empty mel filter 0 (summation-noise column, see test_oracle_features)
    return f


def _tol(prec, scale):
    p95, mx = (2e-3, 2e-3) if prec == 1 else (3e-2, 1.5e-1)
    return p95 * scale, mx * scale


def _check(got, want, prec, what, scale=2.0):
    d = np.abs(got - want)
    p95, mx = _tol(prec, scale)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.percentile(d, 95) <= p95 and d.max() <= mx, (what, float(np.percentile(d, 95)), float(d.max()))


def _oracle_trace(m, enc, t_enc):
    st = DecodeState(m)
    prime(m, st)
    mg = []
    tr = [(t, tok, d) for t, tok, d, _ in tdt_greedy_chunk(m, st, enc, t_enc, margins=mg)]
    amb = next((i for i, (a, b) in enumerate(mg) if a <= TAU or b <= TAU), len(tr))
    return tr, st.tokens, amb


def _check_trace(got, want, amb, prec, what):
    if prec == 1:
        assert got == want, (what, "first difference at", next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want))))
    else:
        assert got[:amb] == want[:amb], (what, amb, next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), -1))


@pytest.mark.parametrize("prec", [1, 0], ids=["precise", "bf16"])
def test_two_utterances_ragged(model_small, oracle_small, features_ref, prec):
    """Two utterances of different length in one call (10 s -> 125 encoder frames, 31 s -> 388: several 64-row tiles and ragged
    tails), features in the contract's bins-major layout; every utterance equals the oracle run on it alone."""
    m = oracle_small
    eng = binding.Engine(model_small, max_streams=2, precision=prec, max_rows=640)
    fs = [_feats(features_ref, 10.0, 1234), _feats(features_ref, 31.0, 77)]
    sids = [eng.open(), eng.open()]
    outs = eng.offline_utterances(sids, features=fs, bins_major=True, want_encoder_output=True, decode=True)
    min_amb = []
    for i, f in enumerate(fs):
        enc, el = m.offline(torch.from_numpy(f[None]), torch.tensor([f.shape[1]]))
        assert outs[i].shape[1] == int(el) == binding.load_library().pkb_encoded_length(f.shape[1])
        _check(outs[i], enc[0].numpy(), prec, f"utterance {i} encoder_output")
        want, toks, amb = _oracle_trace(m, enc, int(el))
        got = eng.last_steps(sids[i])
        _check_trace(got, want, amb, prec, f"utterance {i} trace")
        min_amb.append(amb)
        if prec == 1:
            assert eng.tokens(sids[i]) == toks
    if prec == 0:
        assert max(min_amb) >= 10, "margin filter left nothing to compare"
    eng.close()


@pytest.mark.parametrize("prec", [1, 0], ids=["precise", "bf16"])
def test_config1_full_model_from_audio(model_full, features_ref, prec):
    """BASELINE config 1: one synthetic 10 s 16 kHz clip, batch 1, 24 layers, audio in -> tokens out (GPU log-mel + per-feature
    normalisation + encoder over all 998 frames + TDT).  The oracle is fed the GPU's own features (the frontend has its own
    parity test) so this isolates encoder + decode."""
    m = ModelRef(model_full)
    eng = binding.Engine(model_full, max_streams=1, precision=prec, max_rows=256)
    pcm = synth_clip(10.0, 1234)
    f = eng.logmel(pcm, per_feature_norm=True)          # [T,128]
    assert f.shape == (998, 128)
    sid = eng.open()
    outs = eng.offline_utterances([sid], audio=[pcm], per_feature_norm=True, want_encoder_output=True, decode=True)
    enc, el = m.offline(torch.from_numpy(np.ascontiguousarray(f.T)[None]), torch.tensor([f.shape[0]]))
    assert int(el) == 125 and outs[0].shape == (1024, 125)
    _check(outs[0], enc[0].numpy(), prec, "config 1 encoder_output")
    want, toks, amb = _oracle_trace(m, enc, 125)
    _check_trace(eng.last_steps(sid), want, amb, prec, "config 1 trace")
    if prec == 1:
        assert eng.tokens(sid) == toks
    # a second utterance on the same (reset) stream reproduces the first one exactly
    first = eng.last_steps(sid)
    eng.reset(sid)
    eng.offline_utterances([sid], audio=[pcm], per_feature_norm=True, decode=True)
    assert eng.last_steps(sid) == first
    eng.close()


@pytest.mark.parametrize("prec", [1, 0], ids=["precise", "bf16"])
def test_four_minutes(model_small, oracle_small, features_ref, prec):
    """240 s -> 3000 encoder frames (47 key tiles per query tile, relative positions up to +-2999)."""
    m = oracle_small
    eng = binding.Engine(model_small, max_streams=1, precision=prec, max_rows=3072)
    f = _feats(features_ref, 240.0, 5)
    sid = eng.open()
    outs = eng.offline_utterances([sid], features=[np.ascontiguousarray(f.T)], bins_major=False, want_encoder_output=True, decode=True)
    enc, el = m.offline(torch.from_numpy(f[None]), torch.tensor([f.shape[1]]))
    _check(outs[0], enc[0].numpy(), prec, "4 min encoder_output")
    want, toks, amb = _oracle_trace(m, enc, int(el))
    _check_trace(eng.last_steps(sid), want, amb, prec, "4 min trace")
    eng.close()


def test_long_form_properties(model_small, features_ref):
    """Size-independent properties at a size the oracle cannot reach in seconds (20 min, 15 000 encoder frames, bf16 mode):
    the same clip twice in one batch gives bit-identical outputs (no cross-utterance leakage, deterministic), and a clip
    shorter than one key tile appended to the batch matches its stand-alone result to bf16 tolerance."""
    eng = binding.Engine(model_small, max_streams=3, precision=0, max_rows=2 * 15008 + 64)
    f = _feats(features_ref, 1200.0, 9)
    short = _feats(features_ref, 4.0, 10)
    sids = [eng.open() for _ in range(3)]
    outs = eng.offline_utterances(sids, features=[f, f, short], want_encoder_output=True, decode=True)
    assert outs[0].shape[1] == 15000
    assert np.array_equal(outs[0], outs[1]) and np.isfinite(outs[0]).all()
    assert eng.last_steps(sids[0]) == eng.last_steps(sids[1]) and len(eng.last_steps(sids[0])) >= 15000 // 4
    for s in sids:
        eng.reset(s)
    # (stand-alone, the small GEMMs pick another k-split, so this comparison is to bf16 tolerance, not bit-exact)
    o2 = eng.offline_utterances([sids[0]], features=[short], want_encoder_output=True, decode=False)
    _check(o2[0], outs[2], 0, "short utterance: batched vs alone")
    eng.close()


def test_capacity_error(model_small, features_ref):
    eng = binding.Engine(model_small, max_streams=1, precision=0, max_rows=64)
    sid = eng.open()
    with pytest.raises(RuntimeError, match="row capacity"):
        eng.offline_utterances([sid], features=[_feats(features_ref, 10.0, 1)], decode=False)
    eng.close()


@pytest.mark.parametrize("prec", [1, 0], ids=["precise", "bf16"])
def test_tiny_and_odd_lengths(model_small, oracle_small, features_ref, prec):
    """Edge sizes in one batch: a single feature frame (1 encoder frame), 9 frames (2), 63 / 64 / 65 encoder frames (tile edge)."""
    m = oracle_small
    full = _feats(features_ref, 6.0, 42)
    lens = [1, 9, 8 * 63 - 3, 8 * 64 - 7, 8 * 64 + 1]
    fs = [np.ascontiguousarray(full[:, :n]) for n in lens]
    eng = binding.Engine(model_small, max_streams=len(fs), precision=prec, max_rows=256)
    sids = [eng.open() for _ in fs]
    outs = eng.offline_utterances(sids, features=fs, want_encoder_output=True, decode=True)
    for i, f in enumerate(fs):
        enc, el = m.offline(torch.from_numpy(f[None]), torch.tensor([f.shape[1]]))
        assert outs[i].shape == (1024, int(el)), (lens[i], outs[i].shape, int(el))
        _check(outs[i], enc[0].numpy(), prec, f"{lens[i]} frames")
        want, toks, amb = _oracle_trace(m, enc, int(el))
        _check_trace(eng.last_steps(sids[i]), want, amb, prec, f"{lens[i]} frames trace")
    assert [outs[i].shape[1] for i in range(5)] == [1, 2, 63, 64, 65]
    eng.close()


def test_offline_argument_errors(model_small):
    eng = binding.Engine(model_small, max_streams=2, precision=0, max_rows=64)
    s = eng.open()
    with pytest.raises(RuntimeError, match="at least one feature frame"):
        eng.offline_utterances([s], audio=[np.zeros(399, np.float32)])
    eng.push_features(s, np.zeros((128, 41), np.float32))          # a stream that already streams cannot switch
    eng.step()
    with pytest.raises(RuntimeError, match="freshly opened"):
        eng.offline_utterances([s], features=[np.zeros((128, 40), np.float32)])
    with pytest.raises(RuntimeError, match="bad stream id"):
        eng.offline_utterances([7], features=[np.zeros((128, 40), np.float32)])
    s2 = eng.open()
    with pytest.raises(RuntimeError, match="one utterance per call"):
        eng.offline_utterances([s2, s2], features=[np.zeros((128, 40), np.float32)] * 2)
    eng.close()


@pytest.mark.parametrize("prec", [1, 0], ids=["precise", "bf16"])
def test_deferred_batched_decode(model_small, features_ref, prec):
    """decode=2 parks the encoder rows of utterances encoded in separate calls; pkb_offline_decode_pending decodes them in one
    batched loop.  With the same encoder groups the encoder rows are the same bits, and the per-utterance decode arithmetic does
    not depend on how many utterances share the loop (same joint backend for <= 16), so the traces equal immediate decoding."""
    fs = [_feats(features_ref, sec, 60 + i) for i, sec in enumerate((3.0, 12.0, 7.5, 20.0, 5.0))]
    eng = binding.Engine(model_small, max_streams=5, precision=prec, max_rows=512)
    sids = [eng.open() for _ in fs]
    want = []
    groups = [(0, 2), (2, 3), (3, 5)]
    for lo, hi in groups:                           # immediate decode, same encoder groups (same GEMM shapes -> same encoder bits)
        eng.offline_utterances(sids[lo:hi], features=fs[lo:hi], decode=True)
        for s in sids[lo:hi]:
            want.append((eng.last_steps(s), eng.tokens(s), eng.token_frames(s)))
            eng.reset(s)
    eng.offline_utterances(sids[:2], features=fs[:2], decode=2)      # encoded in three groups ...
    eng.offline_utterances(sids[2:3], features=fs[2:3], decode=2)
    assert eng.tokens(sids[0]) == [] and eng.last_steps(sids[0]) == []
    eng.offline_utterances(sids[3:], features=fs[3:], decode=2)
    assert eng.offline_decode_pending() == 5                         # ... decoded together
    assert eng.offline_decode_pending() == 0
    for s, (steps, toks, frames) in zip(sids, want):
        assert eng.last_steps(s) == steps and eng.tokens(s) == toks and eng.token_frames(s) == frames
    eng.close()


def test_tcgen05_attention_against_the_other_kernels_at_15000_frames(model_small, features_ref):
    """The whole-utterance attention exists three times: lf_attention_tc_kernel (tcgen05 / TMEM, the bf16-mode default),
    lf_attention_tma_kernel (mma.sync; PARAKEET_B200_LF_ATTN=1) and lf_attention_f32_kernel (CUDA cores, precise mode: the kernel
    whose traces equal the oracle's at 125 / 388 / 3000 frames).  The oracle cannot reach 15 000 encoder frames in seconds, so at that
    size (235 key tiles per query tile, 118 query tiles, relative positions up to +-14 999) the tcgen05 kernel is cross-checked
    against both: bf16 tolerance vs the f32 kernel, and the two bf16 kernels agree to bf16 rounding noise of the encoder output."""
    import os
    f = _feats(features_ref, 1200.0, 9)
    outs = {}
    for name, prec, kind in (("tc", 0, "2"), ("tma", 0, "1"), ("f32", 1, None)):
        if kind is not None:
            os.environ["PARAKEET_B200_LF_ATTN"] = kind
        try:
            eng = binding.Engine(model_small, max_streams=1, precision=prec, max_rows=15008 + 64)
            sid = eng.open()
            outs[name] = eng.offline_utterances([sid], features=[f], want_encoder_output=True, decode=True)[0]
            outs[name + "_steps"] = eng.last_steps(sid)
            eng.close()
        finally:
            os.environ.pop("PARAKEET_B200_LF_ATTN", None)
    assert outs["tc"].shape == (1024, 15000) and np.isfinite(outs["tc"]).all()
    _check(outs["tc"], outs["f32"], 0, "tcgen05 (bf16) vs f32 CUDA-core kernel, 15000 frames")
    _check(outs["tma"], outs["f32"], 0, "mma.sync (bf16) vs f32 CUDA-core kernel, 15000 frames")
    d = np.abs(outs["tc"] - outs["tma"])
    assert np.percentile(d, 95) <= 3e-2 and d.max() <= 1.5e-1, (float(np.percentile(d, 95)), float(d.max()))
    # decode traces (diagnostic: a borderline bf16 decision that emits / drops a token shifts everything after it)
    a, b = outs["tc_steps"], outs["f32_steps"]
    n = min(len(a), len(b))
    same = sum(x == y for x, y in zip(a[:n], b[:n]))
    print(f"\n[lf attention 15000 frames] tcgen05 vs f32: max|d| {float(np.abs(outs['tc'] - outs['f32']).max()):.3e}; tcgen05 vs mma.sync: "
          f"max|d| {float(d.max()):.3e}; decode steps equal to the f32 run: {same}/{n}")
    assert n >= 15000 // 4
