"""Kernel variants that must not change results, checked against each other in separate processes (the switches are read once).

* PARAKEET_B200_ATTN_TRIM=0 -- the streaming attention fetches whole 96-key ring blocks instead of only their valid 8-key groups
  (attn_mma.cu).  Invalid slots are masked by select and their V fragments are zeroed, so every trace, token, cache length and
  exported cache must be IDENTICAL, with cache lengths from 0 to saturated in one batch and the rings wrapped.
* PARAKEET_B200_LOGMEL=0 -- the shared-memory Stockham FFT frontend instead of the register-resident one (frontend.cu): two different
  FFT factorizations in f32, equal to 1e-4 (both are tested against the oracle at the north_star tolerance 1e-3 elsewhere).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import binding

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(model, out, env_extra, n_streams=12, n_chunks=108):
    pkg = os.path.dirname(os.path.abspath(binding.__file__))
    env = dict(os.environ, **env_extra)
    run = subprocess.run([sys.executable, os.path.join(HERE, "_ab_driver.py"), pkg, model, out, str(n_streams), str(n_chunks)], env=env,
                         capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and "AB-OK" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]
    return np.load(out)


@pytest.fixture(scope="module")
def default_run(model_small, tmp_path_factory):
    return _run(model_small, str(tmp_path_factory.mktemp("ab") / "default.npz"), {})


def test_attention_valid_slot_loads_equal_whole_block_loads(model_small, default_run, tmp_path):
    whole = _run(model_small, str(tmp_path / "whole.npz"), {"PARAKEET_B200_ATTN_TRIM": "0"})
    a = default_run
    assert a["lens"].max() == 256 and a["lens"].min() == 0 and len(np.unique(a["lens"][60])) >= 8, "mixed cache lengths, saturated at the end"
    assert (a["tokens"] >= 0).sum() > 0
    for k in ("traces", "lens", "tokens", "state0_ch", "state0_tm", "state11_ch", "state11_tm"):
        np.testing.assert_array_equal(a[k], whole[k], err_msg=k)


def test_register_fft_frontend_equals_shared_memory_fft_frontend(model_small, default_run, tmp_path, features_ref):
    from synth_audio import synth_clip
    old = _run(model_small, str(tmp_path / "stockham.npz"), {"PARAKEET_B200_LOGMEL": "0"}, n_streams=2, n_chunks=4)
    for k in ("logmel0", "logmel1", "logmel2"):
        assert default_run[k].shape == old[k].shape and default_run[k].shape[0] > 0
        assert np.max(np.abs(default_run[k] - old[k])) < 1e-4, k
    live = np.ones(128, bool)
    live[0] = False      # empty mel filter 0 (see test_gpu_frontend.test_per_feature_norm)
    assert np.max(np.abs(default_run["logmel_norm"][:, live] - old["logmel_norm"][:, live])) < 1e-3
    # and both against the oracle at the north_star tolerance
    ref = features_ref.logmel(synth_clip(10.0, 1234))
    assert np.max(np.abs(default_run["logmel0"] - ref)) < 1e-3 and np.max(np.abs(old["logmel0"] - ref)) < 1e-3
