"""Known-answer tests recoverable from the reference's checked-in artifacts, run against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, normalized_features
from model_ref import DecodeState, ModelRef, decode_text, prime, streaming_schedule, tdt_greedy_chunk


def tokens_after_subsampling(T):
    for _ in range(3):
        T = (T - 1) // 2 + 1
    return T


def cache_keep(T_in, drop, drop_pre=2):
    return tokens_after_subsampling(T_in) - drop_pre - drop


def test_length_arithmetic_legacy_config():
    # artifacts/diagnostics/streaming_cache_200.jsonl:1,2,100,200: (592, drop 72) keeps 0; (593, 72) keeps 1
    assert cache_keep(592, 72) == 0 and cache_keep(593, 72) == 1
    # docs/VALIDATION_REPORT_TRACE.md:173,177: 48-frame chunks keep 1; 209-213: 41 keeps 1, 57 keeps 3
    assert cache_keep(48, 3) == 1 and cache_keep(41, 3) == 1 and cache_keep(57, 3) == 3
    assert min(64 + cache_keep(48, 3), 256) == 65


def test_schedule_generator_legacy():
    # streaming_cache_200.jsonl `schedule` fields for chunk [592,584], shift [8,8], pre_encode [0,9]
    s = streaming_schedule(200, chunk_size=(592, 584), shift_size=(8, 8), pre_encode=(0, 9))
    assert s[0] == (0, 592) and s[1] == (0, 592) and s[99] == (783, 1376) and s[199] == (1583, 2176)


def test_schedule_generator_current():
    s = streaming_schedule(4)
    assert s == [(0, 41), (8, 65), (32, 89), (56, 113)]
    assert all(e - b == 57 for b, e in s[1:])


def test_streaming_cache_lengths_and_layouts(oracle_small, features_ref):
    m = oracle_small
    feats = torch.from_numpy(normalized_features(features_ref, 4.0, 1234))
    cc, ct, cl = m.initial_cache(1)
    lens = []
    for b, e in streaming_schedule(8):
        enc, el, cc, ct, cl = m.stream_step(feats[None, :, b:e], torch.tensor([e - b]), cc, ct, cl)
        assert enc.shape == (1, 1024, 3) and int(el) == 3                       # encoded_lengths = 3 everywhere (:209-213)
        assert cc.shape == (1, m.L, 256, 1024) and ct.shape == (1, m.L, 1024, 4)
        assert float(ct[..., -1].abs().max()) == 0.0                            # conv cache last column is zero (:212)
        n = int(cl)
        lens.append(n)
        assert float(cc[0, :, : 256 - n].abs().max()) == 0.0                    # unfilled region: zero, valid region = SUFFIX
        assert float(cc[0, :, 256 - n:].abs().min(dim=-1).values.max()) > 0.0
    assert lens == [1, 4, 7, 10, 13, 16, 19, 22]                                # 1 then +3 per chunk


def test_isolated_48_frame_chunks(oracle_small, features_ref):
    # the CLI's --stream-sim 0.5 pushes disjoint 48-frame chunks: cache_len_out 1,2,3,4 (docs/VALIDATION_REPORT_TRACE.md:173)
    m = oracle_small
    feats = torch.from_numpy(normalized_features(features_ref, 2.5, 5))
    cc, ct, cl = m.initial_cache(1)
    lens = []
    for k in range(4):
        _, el, cc, ct, cl = m.stream_step(feats[None, :, 48 * k:48 * k + 48], torch.tensor([48]), cc, ct, cl)
        lens.append(int(cl))
        assert int(el) == 3
    assert lens == [1, 2, 3, 4]
    cl64 = torch.tensor([64])
    *_, out = m.stream_step(feats[None, :, :48], torch.tensor([48]), cc, ct, cl64)
    assert int(out) == 65                                                       # :177


def test_streaming_equals_offline_when_context_is_complete(oracle_small, features_ref):
    """Sanity of the cache semantics: the first chunk (no cache yet, 41 frames) must equal the offline encoder on the
    same 41 frames for the tokens whose whole receptive field is inside the chunk -- except that streaming drops the
    first two pre-encode tokens and zero-pads the conv on the right."""
    m = oracle_small
    feats = torch.from_numpy(normalized_features(features_ref, 1.0, 3))[None, :, :41]
    cc, ct, cl = m.initial_cache(1)
    enc_s, _, _, _, _ = m.stream_step(feats, torch.tensor([41]), cc, ct, cl)
    assert torch.isfinite(enc_s).all() and float(enc_s.abs().max()) < 50


class _FakeModel:
    """argmax oracle of cpp/src/greedy_decode_smoke.cpp:25-37 expressed as joint logits (plain RNNT: tokens carry duration 0,
    blank carries duration 1)."""
    blank, vocab, n_dur = 8192, 8193, 5
    vocab_lines = ["a"] * 8192
    script = {0: [1, 8192], 1: [8192], 2: [2, 3, 8192]}

    def __init__(self):
        self.u = {}

    def joint_logits(self, enc, g):
        t = int(enc[0, 0, 0])
        k = self.u.get(t, 0)
        tok = self.script[t][k]
        self.u[t] = k + 1
        lg = torch.zeros(1, 1, 1, 8198)
        lg[0, 0, 0, tok] = 5.0
        lg[0, 0, 0, 8193 + (1 if tok == 8192 else 0)] = 5.0
        return lg

    def predictor_step(self, y, h, c):
        return torch.zeros(1, 640, 1), h, c

    def is_punct_only(self, tok):
        return False


def test_greedy_control_flow_kat():
    fm = _FakeModel()
    st = DecodeState.__new__(DecodeState)
    st.h = st.c = torch.zeros(2, 1, 640)
    st.g, st.y_id, st.tokens = torch.zeros(1, 640, 1), 8192, []
    enc = torch.arange(3, dtype=torch.float32).view(1, 1, 3)
    tdt_greedy_chunk(fm, st, enc, 3)
    assert st.tokens == [1, 2, 3]
    exe = os.path.join(ROOT, "oracle", "_ref", "greedy_decode_smoke")   # the reference's own smoke binary, when built here
    if os.path.exists(exe):
        assert "[1,2,3]" in subprocess.run([exe], capture_output=True, text=True).stdout


def test_greedy_rules():
    """blank+dur0 advances 1; 8 symbols without advance forces +1; leftover advance is dropped (contract.json:244-253)."""
    class M(_FakeModel):
        def joint_logits(self, enc, g):
            lg = torch.zeros(1, 1, 1, 8198)
            lg[0, 0, 0, 7] = 5.0          # always token 7 with duration 0
            lg[0, 0, 0, 8193] = 5.0
            return lg
    st = DecodeState.__new__(DecodeState)
    st.h = st.c = torch.zeros(2, 1, 640)
    st.g, st.y_id, st.tokens = torch.zeros(1, 640, 1), 8192, []
    tr = tdt_greedy_chunk(M(), st, torch.zeros(1, 1, 2), 2)
    assert len(tr) == 16 and st.tokens == [7] * 16 and [t for t, *_ in tr] == [0] * 8 + [1] * 8

    class B(_FakeModel):
        def joint_logits(self, enc, g):
            lg = torch.zeros(1, 1, 1, 8198)
            lg[0, 0, 0, 8192] = 5.0       # blank with duration 0 -> must advance by 1
            lg[0, 0, 0, 8193] = 5.0
            return lg
    st.tokens = []
    tr = tdt_greedy_chunk(B(), st, torch.zeros(1, 1, 3), 3)
    assert [(t, a) for t, _, _, a in tr] == [(0, 1), (1, 1), (2, 1)] and st.tokens == []

    class D4(_FakeModel):
        def joint_logits(self, enc, g):
            lg = torch.zeros(1, 1, 1, 8198)
            lg[0, 0, 0, 8192] = 5.0
            lg[0, 0, 0, 8193 + 4] = 5.0   # duration 4 from frame 0 of a 3-frame chunk: leftover dropped
            return lg
    tr = tdt_greedy_chunk(D4(), st, torch.zeros(1, 1, 3), 3)
    assert len(tr) == 1


def test_tokenizer_against_reference(model_small):
    """decode_text / is_punct_only vs the reference's own Tokenizer (compiled from /root/reference/cpp/src/tokenizer.cpp into
    oracle/_ref when the tree is present), else vs the committed golden made from it (tests/golden/make_golden.py)."""
    import ctypes
    import json
    m = ModelRef.__new__(ModelRef)
    with open(os.path.join(model_small, "vocab.txt"), encoding="utf-8") as f:
        m.vocab_lines = [ln.rstrip("\n") for ln in f]
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "tokenizer_cases.json")))
    for case in golden["cases"]:
        assert decode_text(golden["vocab"], case["ids"]) == case["text"]
    for i, flag in enumerate(golden["punct_only"]):
        mm = ModelRef.__new__(ModelRef)
        mm.vocab_lines = golden["vocab"]
        assert mm.is_punct_only(i) == bool(flag), golden["vocab"][i]
    so = os.path.join(ROOT, "oracle", "_ref", "libref_tokenizer.so")
    if os.path.exists(so):
        lib = ctypes.CDLL(so)
        lib.reftok_open.restype = ctypes.c_void_p
        lib.reftok_open.argtypes = [ctypes.c_char_p]
        lib.reftok_decode.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        lib.reftok_is_punct_only.argtypes = [ctypes.c_void_p, ctypes.c_int]
        t = lib.reftok_open(os.path.join(model_small, "vocab.txt").encode())
        rng = np.random.default_rng(0)
        for _ in range(50):
            ids = rng.integers(0, 8192, size=int(rng.integers(0, 30))).astype(np.int32)
            buf = ctypes.create_string_buffer(4096)
            lib.reftok_decode(t, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(ids), buf, 4096)
            assert decode_text(m.vocab_lines, ids.tolist()) == buf.value.decode("utf-8")
        for i in range(0, 8192, 7):
            assert m.is_punct_only(i) == bool(lib.reftok_is_punct_only(t, i))


def test_prime_and_decode_runs(oracle_small, features_ref):
    m = oracle_small
    st = DecodeState(m)
    prime(m, st)
    assert st.y_id == m.vocab_lines.index("<|en|>") and st.g.shape == (1, 640, 1)
    feats = torch.from_numpy(normalized_features(features_ref, 2.0, 9))
    cc, ct, cl = m.initial_cache(1)
    n_steps = 0
    for b, e in streaming_schedule(5):
        enc, el, cc, ct, cl = m.stream_step(feats[None, :, b:e], torch.tensor([e - b]), cc, ct, cl)
        tr = tdt_greedy_chunk(m, st, enc, int(el))
        n_steps += len(tr)
        assert all(0 <= t < 3 and 0 <= tok <= 8192 and 0 <= d <= 4 for t, tok, d, _ in tr)
    assert n_steps >= 5


def test_tdt_loop_matches_the_references_own_python_loop(features_ref):
    """tests/golden/tdt_trace_ref.json is the trace of the REFERENCE's tools/verify_nemo/tdt_trace.py, executed unmodified on the
    oracle's encoder / predictor / joint for the seeded 2-layer model (tests/golden/make_tdt_trace_golden.py).  The oracle's own
    loop (tdt_greedy_chunk + prime: the restatement the GPU decode is checked against) must make the same decisions: priming token,
    per-chunk time indices, u counters, blank + duration-0 clamp, duration advance, leftover advance dropped at the chunk end.
    (The Python reference has no leading-punctuation suppression -- that is the C++ runtime's, parakeet_trt.cpp:3256-3262 -- so it is
    switched off here.)  Steps are compared up to the first decision whose top-2 gap is below 1e-4 (CPU BLAS differences)."""
    import json
    import os
    from conftest import model_dir, normalized_features
    from model_ref import DecodeState, ModelRef, prime, tdt_greedy_chunk
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tdt_trace_ref.json")))
    m = ModelRef(model_dir(2))
    f = normalized_features(features_ref, doc["clip"]["seconds"], doc["clip"]["seed"])
    f[0] = 0.0
    st = DecodeState(m)
    prime(m, st)
    assert st.y_id == doc["meta"]["y0"] and doc["meta"]["blank_id"] == m.blank
    got = []
    C = doc["chunk_frames"]
    for ci, lo in enumerate(range(0, f.shape[1], C)):
        seg = f[:, lo:lo + C]
        enc, el = m.offline(torch.from_numpy(np.ascontiguousarray(seg)[None]), torch.tensor([seg.shape[1]]))
        prev_t, u = -1, 0
        for t, tok, d, adv in tdt_greedy_chunk(m, st, enc, int(el), punct_suppression=False):
            u = u + 1 if t == prev_t else 0
            prev_t = t
            got.append((ci, t, u, tok, d, adv))
    want = [(s_["chunk_idx"], s_["time_idx"], s_["u"], s_["best_tok"], s_["duration"], s_["advance"]) for s_ in doc["steps"]]
    amb = next((i for i, s_ in enumerate(doc["steps"]) if min(s_["tok_gap"], s_["dur_gap"]) < 1e-4), len(want))
    assert amb >= 60, amb
    assert got[:amb] == want[:amb]
    if amb == len(want):
        assert len(got) == len(want)


def test_gpu_library_trace_passes_the_references_compare_tool(tmp_path):
    """tests/golden/tdt_steps_b200_stderr.log is the PARAKEET_DEBUG_TDT_STEPS trace this library printed on a B200 for the clip of
    tdt_trace_ref.json (scripts/gpu_trace.py: legacy session, offline encoder mode, fp32-grade arithmetic).  The reference's own
    triage tool tools/verify_nemo/compare_tdt_trace.py, given that log and the reference PyTorch-loop trace, must report a full
    match (it does: "OK: matched 76 steps").  Where the reference tree is absent the same comparison runs on the restated regex."""
    import json
    import os
    import re
    here = os.path.dirname(os.path.abspath(__file__))
    doc = json.load(open(os.path.join(here, "golden", "tdt_trace_ref.json")))
    log = os.path.join(here, "golden", "tdt_steps_b200_stderr.log")
    pt = tmp_path / "pt.jsonl"
    with open(pt, "w") as f:
        f.write(json.dumps(doc["meta"]) + "\n")
        for i, s_ in enumerate(doc["steps"]):
            blank = s_["best_tok"] == doc["meta"]["blank_id"]
            f.write(json.dumps({"type": "step", "step_idx": i, **{k: s_[k] for k in ("chunk_idx", "time_idx", "u", "best_tok", "best_dur_idx", "duration", "advance")},
                                "is_blank": blank, "blank_dur0_clamped": bool(blank and s_["duration"] == 0)}) + "\n")
    tool = "/root/reference/tools/verify_nemo/compare_tdt_trace.py"
    if os.path.exists(tool):
        run = subprocess.run([sys.executable, tool, "--pt-trace", str(pt), "--cpp-stderr", log, "--check-index", "--fields",
                              "best_tok,best_dur_idx,duration,advance,is_blank,blank_dur0_clamped"], capture_output=True, text=True)
        assert run.returncode == 0 and f"matched {len(doc['steps'])} steps" in run.stdout, run.stdout + run.stderr
    rx = re.compile(r"tdt_step time_idx=(\d+) u=(\d+) best_tok=(\d+) best_dur_idx=(\d+) duration=(\d+) advance=(\d+) blank=(\d) blank_dur0_clamped=(\d)")
    got = [tuple(int(x) for x in m_.groups()) for m_ in map(rx.search, open(log)) if m_]
    want = [(s_["time_idx"], s_["u"], s_["best_tok"], s_["best_dur_idx"], s_["duration"], s_["advance"], int(s_["best_tok"] == 8192),
             int(s_["best_tok"] == 8192 and s_["duration"] == 0)) for s_ in doc["steps"]]
    assert got == want
