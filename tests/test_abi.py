"""The C-ABI library loads without a GPU, exports every symbol include/*.h declares, keeps the reference's struct layouts
and error conventions, and fails LOUDLY (no CPU fallback) when asked to compute without a B200."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT

import binding


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:parakeet|trt_asr|pkb)_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(binding.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return binding.load_library()


def test_every_declared_symbol_is_exported(lib):
    names = _declared("parakeet_trt.h") + _declared("trt_asr.h") + _declared("parakeet_b200.h")
    assert set(binding.LEGACY_SYMBOLS) == set(_declared("parakeet_trt.h"))
    assert set(binding.TRT_ASR_SYMBOLS) == set(_declared("trt_asr.h"))
    assert set(binding.B200_SYMBOLS) == set(_declared("parakeet_b200.h"))
    for n in names:
        assert hasattr(lib, n), n


def test_symbol_list_covers_the_reference_mock_library():
    """The reference's -DPARAKEET_MOCK build (cpp/src/mock_lib.cpp + trt_asr.cpp) is its own ABI fixture; ours must export a
    superset (it adds parakeet_set_debug_context, which the mock omits but rust/parakeet_trt/src/lib.rs:69-74 binds)."""
    mock = os.path.join(ROOT, "oracle", "_ref", "libparakeet_trt_mock.so")
    if not os.path.exists(mock):
        pytest.skip("reference mock library not built here (needs /root/reference)")
    def syms(p):
        out = subprocess.run(["nm", "-D", "--defined-only", p], capture_output=True, text=True).stdout
        return {l.split()[-1] for l in out.splitlines() if re.search(r" T (parakeet_|trt_asr_)", l)}
    ref, ours = syms(mock), syms(binding.LIB_PATH)
    assert len(ref) == 11 and ref <= ours and "parakeet_set_debug_context" in ours - ref


def test_struct_layouts():
    assert ctypes.sizeof(binding.ParakeetConfig) == 16          # char* + int32 + bool (+pad), as bindgen sees the reference header
    assert ctypes.sizeof(binding.ParakeetEvent) == 24
    assert binding.ParakeetEvent.text.offset == 8 and binding.ParakeetEvent.error_message.offset == 16
    assert ctypes.sizeof(binding.TrtAsrEvent) == 32 and binding.TrtAsrEvent.text.offset == 16


def test_null_argument_conventions(lib):
    assert lib.parakeet_create_session(None) is None                                   # parakeet_trt.cpp:1701
    cfg = binding.ParakeetConfig(None, 0, True)
    assert lib.parakeet_create_session(ctypes.byref(cfg)) is None
    lib.parakeet_destroy_session(None)                                                 # :1846 NULL-safe
    lib.parakeet_reset_utterance(None)
    lib.parakeet_set_debug_context(None, b"x", 0, 0, 0)
    x = np.zeros(128, np.float32)
    assert lib.parakeet_push_features(None, x.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 1) == -1   # :1968
    assert lib.parakeet_poll_event(None, None) is False                               # :3861
    assert lib.trt_asr_create_session(None) is None
    assert lib.trt_asr_push_features_f32(None, None, 0, 0) == -1
    assert lib.trt_asr_poll_event(None, None) is False
    assert lib.pkb_engine_create(None) is None and b"null" in lib.pkb_last_error()
    assert lib.pkb_stream_open(None) == -1


def test_missing_model_dir_fails_loudly(lib, capfd):
    cfg = binding.ParakeetConfig(b"/nonexistent/model", 0, True)
    assert lib.parakeet_create_session(ctypes.byref(cfg)) is None
    assert "create_session failed" in capfd.readouterr().err                            # message to stderr (:1839-1842)


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib, model_small):
    with pytest.raises(RuntimeError):
        binding.Engine(model_small, max_streams=1)
    with pytest.raises(RuntimeError):
        binding.ParakeetSessionSafe(model_small)


def test_weights_container_roundtrip(tmp_path):
    from weights_io import DT_BF16, DT_F32, bf16_bits_to_f32, f32_to_bf16_bits, read_weights, write_weights
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((5, 7)).astype(np.float32), rng.standard_normal(11).astype(np.float32)
    write_weights(str(tmp_path / "w.bin"), {"n_layers": 3, "x": -5}, {"a.weight": (a, DT_BF16), "b": (b, DT_F32)})
    cfg, t = read_weights(str(tmp_path / "w.bin"))
    assert cfg == {"n_layers": 3, "x": -5}
    assert np.array_equal(t["b"], b)
    assert np.array_equal(t["a.weight"], bf16_bits_to_f32(f32_to_bf16_bits(a)).reshape(5, 7))
    assert np.max(np.abs(t["a.weight"] - a)) <= np.max(np.abs(a)) * 2 ** -8
    # round-to-nearest-even at the bf16 boundary
    x = np.array([1.0 + 2 ** -8, 1.0 + 3 * 2 ** -8], np.float32)
    assert f32_to_bf16_bits(x).tolist() == [0x3F80, 0x3F82]


def test_encoded_length_formula():
    """pkb_encoded_length == NeMo calc_length applied three times (floor((L - 1) / 2) + 1); pure host function, no GPU."""
    lib = binding.load_library()
    for n, want in [(0, 0), (1, 1), (2, 1), (9, 2), (41, 6), (57, 8), (256, 32), (998, 125), (359998, 45000)]:
        assert lib.pkb_encoded_length(n) == want, n
