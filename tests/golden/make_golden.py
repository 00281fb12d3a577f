#!/usr/bin/env python3
"""Generates tests/golden/tokenizer_cases.json by running the REFERENCE's own Tokenizer
(/root/reference/cpp/src/tokenizer.cpp, compiled by oracle/Makefile into oracle/_ref/libref_tokenizer.so).
Run in the build container (the reference tree is not on the GPU box); the JSON is committed."""
import ctypes
import json
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VOCAB = ["<unk>", "<pad>", "<|en|>", "<blank>", ".", ",", "▁?", "▁", "▁hello", "wor", "ld", "▁a1", "!!", "▁-", "x.", " ", "▁ ",
         "<", ">", "<>", "a>", "▁the", "re", "'", "▁'s", "é", "▁é", "。", "12", "▁3"]

lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_tokenizer.so"))
lib.reftok_open.restype = ctypes.c_void_p
lib.reftok_open.argtypes = [ctypes.c_char_p]
lib.reftok_decode.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
lib.reftok_is_punct_only.argtypes = [ctypes.c_void_p, ctypes.c_int]
with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, encoding="utf-8") as f:
    f.write("\n".join(VOCAB) + "\n")
t = lib.reftok_open(f.name.encode())
rng = np.random.default_rng(1)
cases = []
seqs = [[8, 9, 10], [4, 8], [7, 8, 7, 9], [0, 1, 2, 3], [], [99, -1, 8], [21, 22, 24, 5, 11], [26, 25, 27], [15, 8, 16, 9]]
seqs += [rng.integers(0, len(VOCAB), size=int(rng.integers(1, 12))).tolist() for _ in range(40)]
for ids in seqs:
    a = np.asarray(ids, np.int32)
    buf = ctypes.create_string_buffer(4096)
    lib.reftok_decode(t, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(ids), buf, 4096)
    cases.append({"ids": [int(i) for i in ids], "text": buf.value.decode("utf-8")})
punct = [int(lib.reftok_is_punct_only(t, i)) for i in range(len(VOCAB))]
json.dump({"vocab": VOCAB, "cases": cases, "punct_only": punct, "source": "reference cpp/src/tokenizer.cpp via oracle/ref_shims"},
          open(os.path.join(ROOT, "tests", "golden", "tokenizer_cases.json"), "w"), ensure_ascii=False, indent=0)
print("wrote", len(cases), "cases")
