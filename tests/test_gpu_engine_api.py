"""Engine-level behaviour around the hot path (through the C ABI): queue bookkeeping, error paths that must not leak stream
slots, the running-statistics normalisation mode, and a converted synthetic .nemo loaded into the GPU engine."""
import io
import os
import tarfile

import numpy as np
import pytest
import torch

import binding
from conftest import normalized_features
from model_ref import DecodeState, ModelRef, prime, streaming_schedule, tdt_greedy_chunk
from synth_audio import synth_clip

pytestmark = pytest.mark.gpu


def _feats(features_ref, seconds, seed):
    f = normalized_features(features_ref, seconds, seed)
    f[0] = 0.0
    return f


def test_feature_ring_accounting_past_one_wrap(model_small, features_ref):
    """the ring-full check compared an absolute frame count with a ring index -- after 512 frames every second queued
    chunk was refused.  Two chunks may wait in the ring at any point of a long stream; a third 256-frame push must be refused."""
    f = _feats(features_ref, 14.0, 3)
    eng = binding.Engine(model_small, max_streams=1, precision=1)
    s = eng.open()
    pos = 0
    for _ in range(12):                      # 12 x 2 x 57 = 1368 frames: the 512-frame ring wraps more than twice
        eng.push_features(s, f[:, pos:pos + 57])
        eng.push_features(s, f[:, pos + 57:pos + 114])
        assert eng.step() == 1 and eng.step() == 1 and eng.step() == 0
        pos += 114
    eng.push_features(s, f[:, :256])
    eng.push_features(s, f[:, :256])
    with pytest.raises(RuntimeError, match="feature ring full"):
        eng.push_features(s, f[:, :256])
    assert eng.step() == 1 and eng.step() == 1
    eng.push_features(s, f[:, :256])         # room again
    assert eng.step() == 1
    eng.close()


def test_has_pending_sees_chunks_already_in_the_ring(model_small):
    """one frontend pass can produce frames for several scheduled chunks; `while has_pending: step()` must drain them."""
    eng = binding.Engine(model_small, max_streams=1, precision=1)
    s = eng.open()
    pcm = synth_clip(2.0, 17)
    eng.push_audio(s, pcm[:8192 * 3])        # 24576 samples = 151 frames = chunks [0,41) [8,65) [32,89) [56,113) [80,137)
    steps = 0
    while eng.has_pending(s):
        n = eng.step()
        steps += n
        assert steps < 50
    assert steps == 5 and eng.chunks_done(s) == 5 and not eng.has_pending(s)
    eng.close()


def test_failed_calls_do_not_leak_stream_slots(model_small, features_ref):
    """tensor-level calls borrow stream slots; a rejected call (bad T, bad cache length, bad length) must give them back."""
    eng = binding.Engine(model_small, max_streams=2, precision=1)
    L = eng.n_layers
    cc, ct = np.zeros((1, L, 256, 1024), np.float32), np.zeros((1, L, 1024, 4), np.float32)
    x8, x57 = np.zeros((1, 128, 8), np.float32), np.zeros((1, 128, 57), np.float32)
    for _ in range(5):                       # more failures than slots
        with pytest.raises(RuntimeError, match="33..256"):
            eng.encoder_streaming_step(x8, np.array([8]), cc, ct, np.array([0]))
        with pytest.raises(RuntimeError, match="out of range"):
            eng.encoder_streaming_step(x57, np.array([57]), cc, ct, np.array([300]))
        with pytest.raises(RuntimeError, match="length must equal T"):
            eng.encoder_streaming_step(x57, np.array([50]), cc, ct, np.array([0]))
        with pytest.raises(RuntimeError, match="length must equal T"):
            eng.encoder_offline_step(x57, np.array([50]))
    enc, el, *_ = eng.encoder_streaming_step(x57, np.array([57]), cc, ct, np.array([0]))      # both slots still free
    assert el.tolist() == [3]
    a, b = eng.open(), eng.open()
    with pytest.raises(RuntimeError, match="no free stream slot"):
        eng.open()
    # a chunk the encoder cannot take is refused when it is pushed and leaves nothing queued
    with pytest.raises(RuntimeError, match="33..256"):
        eng.push_features(a, np.zeros((128, 20), np.float32))
    assert not eng.has_pending(a) and eng.step() == 0
    eng.push_features(a, _feats(features_ref, 1.0, 1)[:, :41])
    assert eng.step() == 1
    eng.close()


def test_running_statistics_normalisation(model_small, features_ref):
    """pkb_stream_set_feature_norm_running: every frame normalised with the causal running mean / unbiased std of its own stream.
    Checked through the decode trace: a stream in running mode fed raw audio == a plain stream fed the same audio's features
    normalised on the host with the numpy restatement below (pushed as explicit chunks cut by the same schedule)."""
    pcm = synth_clip(4.0, 23)
    raw = features_ref.logmel(pcm).astype(np.float64)                       # [T,128]
    T = raw.shape[0]
    n = np.arange(1, T + 1)[:, None]
    mean = np.cumsum(raw, 0) / n
    want = np.zeros_like(raw)
    for t in range(1, T):                                                   # frame 0 has no variance yet -> 0
        var = ((raw[:t + 1] - mean[t]) ** 2).sum(0) / t
        want[t] = (raw[t] - mean[t]) / (np.sqrt(var) + 1e-5)
    want = want.astype(np.float32)
    want[:, 0] = 0.0        # the empty mel filter 0 is the constant ln(1e-5): exactly 0 in both (x - mean == 0)
    eng = binding.Engine(model_small, max_streams=2, precision=1)
    a, b = eng.open(), eng.open()
    eng.set_feature_norm_running(a, True)
    n_chunks = 12
    sched = streaming_schedule(n_chunks)
    # audio stream: 8192-sample pushes (the frames arrive in uneven groups, the statistics must not care)
    for pos in range(0, pcm.size, 8192):
        eng.push_audio(a, pcm[pos:pos + 8192])
    tr_a = []
    while eng.has_pending(a):
        if eng.step():
            tr_a.append(eng.last_steps(a))
    tr_b = []
    for lo, hi in sched:
        eng.push_features(b, np.ascontiguousarray(want[lo:hi].T))
        assert eng.step() == 1
        tr_b.append(eng.last_steps(b))
    assert len(tr_a) >= n_chunks and tr_a[:n_chunks] == tr_b and sum(len(x) for x in tr_b) >= n_chunks
    # the mode survives a reset, the statistics restart: the same audio gives the same trace again
    eng.reset(a)
    eng.push_audio(a, pcm[:8192 * 2])
    again = []
    while eng.has_pending(a):
        if eng.step():
            again.append(eng.last_steps(a))
    assert again and again == tr_a[:len(again)]
    eng.close()


def test_converted_nemo_archive_runs_on_the_gpu(tmp_path, model_small, oracle_small, features_ref):
    """tools/convert_nemo.py end to end into the engine: the 2-layer synthetic model is packed as a .nemo archive of the published
    structure (tar.gz: model_config.yaml + model_weights.ckpt state_dict + tokenizer vocab + junk keys NeMo carries), converted, loaded
    by pkb_engine_create, and must decode exactly like the oracle on the original weights (fp32-grade mode)."""
    from convert_nemo import convert
    from weights_io import read_weights
    cfg, w = read_weights(os.path.join(model_small, "weights.bin"))
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}
    sd["encoder.layers.0.conv.batch_norm.num_batches_tracked"] = torch.tensor(3)
    sd["preprocessor.featurizer.window"] = torch.zeros(400)
    buf = io.BytesIO()
    torch.save(sd, buf)
    vocab = open(os.path.join(model_small, "vocab.txt"), encoding="utf-8").read()
    yaml_txt = f"encoder:\n  n_layers: {cfg['n_layers']}\n  d_model: {cfg['d_model']}\nmodel_defaults:\n  tdt_durations: [0, 1, 2, 3, 4]\n"
    path = str(tmp_path / "m.nemo")
    with tarfile.open(path, "w:gz", compresslevel=1) as tar:
        for name, data in (("./model_config.yaml", yaml_txt.encode()), ("./model_weights.ckpt", buf.getvalue()), ("./abc_vocab.txt", vocab.encode())):
            info = tarfile.TarInfo(name)
            info.size = len(data)
            tar.addfile(info, io.BytesIO(data))
    out = str(tmp_path / "converted")
    convert(path, out)
    eng = binding.Engine(out, max_streams=1, precision=1)
    s = eng.open()
    m = oracle_small
    f = _feats(features_ref, 3.0, 61)
    st = DecodeState(m)
    prime(m, st)
    cc, ct, cl = m.initial_cache(1)
    for b, e in streaming_schedule(8):
        eng.push_features(s, f[:, b:e])
        assert eng.step() == 1
        enc, el, cc, ct, cl = m.stream_step(torch.from_numpy(f[None, :, b:e]), torch.tensor([e - b]), cc, ct, cl)
        want = [(t, tok, d) for t, tok, d, _ in tdt_greedy_chunk(m, st, enc, int(el))]
        assert eng.last_steps(s) == want
    assert eng.tokens(s) == st.tokens and eng.text(s)
    eng.close()


@pytest.mark.parametrize("precision,n_streams", [(1, 1), (0, 24)], ids=["precise-1-stream", "bf16-24-streams"])
def test_step_graph_equals_launch_by_launch(model_small, features_ref, precision, n_streams):
    """Steps of a repeated shape replay ONE CUDA graph (encoder chunk + device-side WHILE node around the decode iteration).  The graph
    must be built (pkb_engine_graphs_built), must count its kernels, and must produce exactly the traces, tokens, cache lengths and
    exported state of the launch-by-launch path (PARAKEET_B200_GRAPH=0), including chunks whose decode loop runs many passes."""
    import os
    n_chunks = 20
    feats = [_feats(features_ref, 0.41 + 0.24 * n_chunks + 0.5, 300 + i) for i in range(min(n_streams, 4))]

    def run(graph: bool):
        if not graph:
            os.environ["PARAKEET_B200_GRAPH"] = "0"
        try:
            eng = binding.Engine(model_small, max_streams=n_streams, precision=precision)
        finally:
            os.environ.pop("PARAKEET_B200_GRAPH", None)
        sids = [eng.open() for _ in range(n_streams)]
        traces, launches = [], []
        for b, e in streaming_schedule(n_chunks):
            for i, s in enumerate(sids):
                eng.push_features(s, feats[i % len(feats)][:, b:e])
            l0 = eng.kernel_launches()
            assert eng.step() == n_streams
            launches.append(eng.kernel_launches() - l0)
            traces.append([eng.last_steps(s) for s in sids])
        out = dict(traces=traces, tokens=[eng.tokens(s) for s in sids], lens=[eng.cache_len(s) for s in sids], graphs=eng.graphs_built(),
                   launches=launches, state=eng.export_state(sids[0]), loop=eng.decode_loop_stats())
        eng.close()
        return out
    g, p = run(True), run(False)
    assert p["graphs"] == 0 and g["graphs"] >= 1, "the steady-state shape must have been captured"
    assert g["traces"] == p["traces"] and g["tokens"] == p["tokens"] and g["lens"] == p["lens"]
    assert sum(len(t) for t in g["tokens"]) > 0
    np.testing.assert_array_equal(g["state"][0], p["state"][0])
    np.testing.assert_array_equal(g["state"][1], p["state"][1])
    # launch accounting: a replayed step reports the kernels its graph ran (same kernels as the launch-by-launch step, minus the
    # one speculative iteration the host-polled loop enqueues after the batch has finished)
    assert all(x > 0 for x in g["launches"])
    steady = [(a, b) for a, b in zip(g["launches"], p["launches"])][3:]
    assert all(0 <= b - a <= 12 for a, b in steady), steady
    ms, nbytes, passes, loops = g["loop"]
    assert loops >= n_chunks - 3 and passes >= loops and ms > 0 and nbytes > 0
