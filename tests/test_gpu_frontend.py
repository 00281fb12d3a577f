"""GPU log-mel frontend (through the C ABI) vs the C restatement of rust/features.  Tolerance: 1e-3 abs (north_star)."""
import numpy as np
import pytest

import binding
from synth_audio import synth_clip

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def eng(model_small):
    e = binding.Engine(model_small, max_streams=4, precision=1)
    yield e
    e.close()


@pytest.mark.parametrize("seconds,seed", [(10.0, 1234), (0.5, 1), (1.0, 2), (3.37, 3)])
def test_logmel_matches_oracle(eng, features_ref, seconds, seed):
    pcm = synth_clip(seconds, seed)
    ref = features_ref.logmel(pcm)
    got = eng.logmel(pcm)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < TOL


def test_edge_cases(eng, features_ref):
    z = eng.logmel(np.zeros(16000, np.float32))
    assert z.shape == (98, 128) and np.all(z == np.float32(np.log(np.float32(1e-5))))      # lib.rs:229-241 + analytic KAT
    assert eng.logmel(np.zeros(399, np.float32)).shape[0] == 0                             # shorter than one window
    assert eng.logmel(np.zeros(0, np.float32)).shape[0] == 0                               # empty input (lib.rs:67-69)
    one = synth_clip(0.025, 5)                                                             # exactly one frame
    assert np.max(np.abs(eng.logmel(one) - features_ref.logmel(one))) < TOL
    loud = np.clip(synth_clip(1.0, 6) * 20, -1, 1)                                         # clipped full-scale input
    assert np.max(np.abs(eng.logmel(loud) - features_ref.logmel(loud))) < TOL
    for n in (559, 560, 561, 16001):                                                       # ragged lengths around the hop
        a = synth_clip(2.0, 8)[:n]
        assert eng.logmel(a).shape == features_ref.logmel(a).shape


def test_per_feature_norm(eng, features_ref):
    pcm = synth_clip(10.0, 1234)
    ref = features_ref.normalized(features_ref.logmel(pcm))
    got = eng.logmel(pcm, per_feature_norm=True)
    live = np.ones(128, bool)
    live[0] = False        # empty mel filter 0: value is summation noise / 1e-5 in the reference too (see test_oracle_features)
    assert np.max(np.abs(got[:, live] - ref[:, live])) < TOL
    # feature 0: both sides hold a constant column (std floor), sign/magnitude of the noise term is order-dependent
    assert np.ptp(got[:, 0]) == 0.0 and np.ptp(ref[:, 0]) == 0.0 and abs(got[0, 0]) < 2.0


def test_linearity_property_at_scale(eng):
    """size-independent property at a BASELINE-sized input (60 s): scaling the audio by 2 adds ln 4 to every live feature."""
    pcm = 0.25 * synth_clip(60.0, 77)
    a, b = eng.logmel(pcm), eng.logmel(2 * pcm)
    assert a.shape == (5998, 128)
    loud = a > -4.0         # where the mel energy dwarfs the 1e-5 floor inside ln(E + 1e-5)
    assert loud.mean() > 0.5
    assert np.max(np.abs((b - a)[loud] - np.log(4.0))) < 1e-3
