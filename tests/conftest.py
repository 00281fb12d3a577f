"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): oracle vs known answers, host logic, ABI symbol checks.
`-m gpu` runs on a B200 and compares the CUDA path (through the C ABI) with the oracle."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "trt-asr-engine_b200")
for p in (ROOT, PKG, os.path.join(PKG, "tools"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class FeaturesRef:
    """ctypes view of oracle/_build/libfeatures_ref.so (the C restatement of rust/features)."""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        self.lib.fr_num_frames.restype = ctypes.c_size_t
        self.lib.fr_num_frames.argtypes = [ctypes.c_size_t]

    def num_frames(self, n):
        return int(self.lib.fr_num_frames(n))

    def logmel(self, audio, threads=1):
        a = np.ascontiguousarray(audio, np.float32)
        T = self.num_frames(a.size)
        out = np.zeros((T, 128), np.float32)
        if T:
            self.lib.fr_logmel(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.size), out.ctypes.data_as(ctypes.c_void_p),
                               ctypes.c_int(threads))
        return out

    def stats(self, feat):
        f = np.ascontiguousarray(feat, np.float32)
        mean, std = np.zeros(128, np.float32), np.zeros(128, np.float32)
        self.lib.fr_per_feature_stats(f.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(f.shape[0]),
                                      mean.ctypes.data_as(ctypes.c_void_p), std.ctypes.data_as(ctypes.c_void_p))
        return mean, std

    def normalized(self, feat):
        f = np.array(feat, np.float32, copy=True, order="C")
        mean, std = self.stats(f)
        self.lib.fr_apply_norm(f.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(f.shape[0]),
                               mean.ctypes.data_as(ctypes.c_void_p), std.ctypes.data_as(ctypes.c_void_p))
        return f

    def tables(self):
        w, fb = np.zeros(400, np.float32), np.zeros((128, 257), np.float32)
        self.lib.fr_get_tables(w.ctypes.data_as(ctypes.c_void_p), fb.ctypes.data_as(ctypes.c_void_p))
        return w, fb


def build_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "libfeatures_ref.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(ROOT, "oracle", "features_ref.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libfeatures_ref.so"], stdout=subprocess.DEVNULL)
    return so


@pytest.fixture(scope="session")
def features_ref():
    return FeaturesRef(build_oracle())


def model_dir(n_layers: int) -> str:
    from make_synthetic_model import ensure_model
    return ensure_model(os.path.join(ROOT, "models", f"synth{n_layers}"), n_layers=n_layers, seed=0)


@pytest.fixture(scope="session")
def model_small():
    return model_dir(2)


@pytest.fixture(scope="session")
def model_full():
    return model_dir(24)


@pytest.fixture(scope="session")
def oracle_small(model_small):
    from model_ref import ModelRef
    return ModelRef(model_small)


def normalized_features(features_ref, seconds, seed):
    """per-feature-normalised features of a synthetic clip, bins-major [128,T] (the C-ABI layout)."""
    from synth_audio import synth_clip
    f = features_ref.normalized(features_ref.logmel(synth_clip(seconds, seed)))
    return np.ascontiguousarray(f.T)
