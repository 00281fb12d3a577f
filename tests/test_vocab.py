"""The PRODUCT's token table (csrc/vocab.cpp behind pkb_vocab_* / pkb_detokenize) against the reference's Tokenizer
(/root/reference/cpp/src/tokenizer.cpp:9-84): the committed golden cases made from the compiled reference
(tests/golden/make_golden.py -> tokenizer_cases.json) and, where oracle/_ref/libref_tokenizer.so exists (the build container), the
compiled reference itself on random id sequences over the synthetic 8192-piece vocabulary.  No GPU involved."""
import ctypes
import json
import os

import numpy as np
import pytest

import binding
from conftest import ROOT

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "tokenizer_cases.json"), encoding="utf-8"))


def _write_vocab(path, pieces):
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(pieces) + "\n")
    return str(path)


def test_product_vocab_matches_the_reference_golden(tmp_path):
    v = binding.Vocab(_write_vocab(tmp_path / "vocab.txt", GOLDEN["vocab"]))
    assert v.vocab_size() == len(GOLDEN["vocab"])
    for case in GOLDEN["cases"]:
        assert v.decode(case["ids"]) == case["text"], case
    assert [int(v.is_punct_only(i)) for i in range(len(GOLDEN["vocab"]))] == GOLDEN["punct_only"]
    assert not v.is_punct_only(-1) and not v.is_punct_only(len(GOLDEN["vocab"]))      # out of range: false (tokenizer.cpp:60)
    v.close()


def test_product_vocab_load_rules(tmp_path):
    lib = binding.load_library()
    assert not lib.pkb_vocab_open(str(tmp_path / "missing.txt").encode()) and b"cannot open" in lib.pkb_last_error()
    (tmp_path / "empty.txt").write_text("")
    assert not lib.pkb_vocab_open(str(tmp_path / "empty.txt").encode()) and b"empty" in lib.pkb_last_error()     # tokenizer.cpp:20-22
    (tmp_path / "crlf.txt").write_bytes("▁a\r\nb\r\n<pad>\r\n".encode("utf-8"))                                      # CR stripped (:17)
    v = binding.Vocab(str(tmp_path / "crlf.txt"))
    assert v.vocab_size() == 3 and v.decode([0, 1, 2, 0]) == "ab a"
    # truncating copy: returns the full length, writes cap-1 bytes + NUL
    ids = np.array([0, 1, 0, 1], np.int32)
    buf = ctypes.create_string_buffer(4)
    n = lib.pkb_vocab_decode(v._v, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 4, buf, 4)
    assert n == len("ab ab") and buf.value == b"ab "


def test_product_vocab_against_the_compiled_reference(model_small):
    so = os.path.join(ROOT, "oracle", "_ref", "libref_tokenizer.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_tokenizer.so is built only where /root/reference exists")
    ref = ctypes.CDLL(so)
    ref.reftok_open.restype = ctypes.c_void_p
    ref.reftok_open.argtypes = [ctypes.c_char_p]
    ref.reftok_decode.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    ref.reftok_is_punct_only.argtypes = [ctypes.c_void_p, ctypes.c_int]
    path = os.path.join(model_small, "vocab.txt")
    t = ref.reftok_open(path.encode())
    v = binding.Vocab(path)
    rng = np.random.default_rng(0)
    for _ in range(200):
        ids = rng.integers(-2, 8200, size=int(rng.integers(0, 40))).astype(np.int32)
        buf = ctypes.create_string_buffer(8192)
        ref.reftok_decode(t, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(ids), buf, 8192)
        assert v.decode(ids) == buf.value.decode("utf-8")
    assert all(v.is_punct_only(i) == bool(ref.reftok_is_punct_only(t, i)) for i in range(-1, 8194))


@pytest.mark.gpu
def test_engine_detokenize_matches_the_reference_golden(tmp_path, model_small):
    """pkb_detokenize / pkb_stream_text / the punctuation predicate of a LIVE engine whose model directory carries the golden vocabulary."""
    d = tmp_path / "m"
    d.mkdir()
    os.symlink(os.path.join(model_small, "weights.bin"), d / "weights.bin")
    _write_vocab(d / "vocab.txt", GOLDEN["vocab"])
    eng = binding.Engine(str(d), max_streams=1)
    for case in GOLDEN["cases"]:
        assert eng.detokenize(case["ids"]) == case["text"], case
    assert [int(eng.token_is_punct_only(i)) for i in range(len(GOLDEN["vocab"]))] == GOLDEN["punct_only"]
    eng.close()
