"""Pins the oracle's encoder restatement against an INDEPENDENT public implementation of the same architecture:
`transformers.models.parakeet.ParakeetEncoder` (Hugging Face's port of NeMo's FastConformer encoder, validated by its authors
against NeMo; transformers 5.5 ships in this image).  NeMo itself, the .nemo weights and the reference's golden JSONL are not
available offline, so this is the strongest numeric pin the sandbox allows for `oracle/model_ref.py`:
the seeded synthetic weights are loaded into both, name by name, and the full-context (`offline`) encoder outputs compared.
What it covers: dw_striding subsampling (conv indices, channel-major flatten), the RelPositionalEncoding table and rel_shift
indexing, pos_bias_u / pos_bias_v, the (ac + bd)/sqrt(d_k) scaling, the macaron FFN halves, GLU -> depthwise -> BatchNorm(eval)
-> SiLU ordering, LayerNorm placement.  The cache-aware streaming step shares `_layer()` with `offline()` and is pinned
structurally in test_oracle_kats.py.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import model_dir, normalized_features
from model_ref import ModelRef

parakeet = pytest.importorskip("transformers.models.parakeet.modeling_parakeet")
from transformers.models.parakeet.configuration_parakeet import ParakeetEncoderConfig      # noqa: E402


def _hf_encoder(m: ModelRef):
    cfg = ParakeetEncoderConfig(hidden_size=m.D, num_hidden_layers=m.L, num_attention_heads=m.H, intermediate_size=m.cfg["ff_dim"],
                                hidden_act="silu", attention_bias=False, convolution_bias=False, conv_kernel_size=m.cfg["conv_kernel"],
                                subsampling_factor=8, subsampling_conv_channels=m.cfg["sub_channels"], num_mel_bins=m.cfg["feat_in"],
                                dropout=0.0, dropout_positions=0.0, layerdrop=0.0, activation_dropout=0.0, attention_dropout=0.0,
                                max_position_embeddings=5000, scale_input=False)      # xscaling false: audit_model_arch.json:34
    cfg._attn_implementation = "eager"
    enc = parakeet.ParakeetEncoder(cfg).eval()
    sd = {}
    w = m.w
    for i in (0, 2, 3, 5, 6):
        sd[f"subsampling.layers.{i}.weight"] = w[f"encoder.pre_encode.conv.{i}.weight"]
        sd[f"subsampling.layers.{i}.bias"] = w[f"encoder.pre_encode.conv.{i}.bias"]
    sd["subsampling.linear.weight"] = w["encoder.pre_encode.out.weight"]
    sd["subsampling.linear.bias"] = w["encoder.pre_encode.out.bias"]
    for l in range(m.L):
        p, q = f"encoder.layers.{l}.", f"layers.{l}."
        for n in ("norm_feed_forward1", "norm_self_att", "norm_conv", "norm_feed_forward2", "norm_out"):
            sd[q + n + ".weight"] = w[p + n + ".weight"]
            sd[q + n + ".bias"] = w[p + n + ".bias"]
        for ff in ("feed_forward1", "feed_forward2"):
            sd[q + ff + ".linear1.weight"] = w[p + ff + ".linear1.weight"]
            sd[q + ff + ".linear2.weight"] = w[p + ff + ".linear2.weight"]
        for a, b in (("linear_q", "q_proj"), ("linear_k", "k_proj"), ("linear_v", "v_proj"), ("linear_out", "o_proj"),
                     ("linear_pos", "relative_k_proj")):
            sd[q + f"self_attn.{b}.weight"] = w[p + f"self_attn.{a}.weight"]
        sd[q + "self_attn.bias_u"] = w[p + "self_attn.pos_bias_u"].reshape(m.H, m.dk)
        sd[q + "self_attn.bias_v"] = w[p + "self_attn.pos_bias_v"].reshape(m.H, m.dk)
        for n in ("pointwise_conv1", "pointwise_conv2"):
            t = w[p + f"conv.{n}.weight"]
            sd[q + f"conv.{n}.weight"] = t if t.dim() == 3 else t.unsqueeze(-1)
        sd[q + "conv.depthwise_conv.weight"] = w[p + "conv.depthwise_conv.weight"]
        for a, b in (("weight", "weight"), ("bias", "bias"), ("running_mean", "running_mean"), ("running_var", "running_var")):
            sd[q + f"conv.norm.{b}"] = w[p + f"conv.batch_norm.{a}"]
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("num_batches_tracked") or k.endswith("inv_freq") for k in missing), missing
    return enc


@pytest.mark.parametrize("seconds", [2.0, 10.0], ids=["2s", "10s"])
def test_offline_encoder_matches_hf_parakeet(features_ref, seconds):
    m = ModelRef(model_dir(2))
    enc = _hf_encoder(m)
    f = normalized_features(features_ref, seconds, 1234)              # [128, T]
    f[0] = 0.0
    x = torch.from_numpy(f[None])
    want, el = m.offline(x, torch.tensor([f.shape[1]]))               # [1, 1024, T_enc]
    with torch.no_grad():
        got = enc(input_features=x.transpose(1, 2)).last_hidden_state  # [1, T_enc, 1024]
    assert got.shape[1] == int(el) == want.shape[2]
    d = (got.transpose(1, 2) - want).abs()
    # two fp32 implementations with different operation order (fused vs separate scaling, conv via Conv1d); measured 1.7e-6 on O(1) outputs
    assert float(d.max()) < 2e-5, float(d.max())
    assert float(want.abs().mean()) > 0.1


def test_predictor_matches_torch_lstm():
    """RNNTDecoder.predict is an embedding + torch.nn.LSTM (NeMo rnnt_modules / rnn.py LSTMDropout): the oracle's hand-written
    cell must equal nn.LSTM loaded with the same weights, step by step with carried state."""
    m = ModelRef(model_dir(2))
    H, L = m.cfg["pred_hidden"], m.cfg["pred_layers"]
    lstm = torch.nn.LSTM(H, H, num_layers=L).eval()
    pre = "decoder.prediction.dec_rnn.lstm."
    lstm.load_state_dict({k: m.w[pre + k] for k in lstm.state_dict()})
    emb = m.w["decoder.prediction.embed.weight"]
    torch.manual_seed(3)
    h = 0.3 * torch.randn(L, 2, H)
    c = 0.3 * torch.randn(L, 2, H)
    h2, c2 = h.clone(), c.clone()
    for y in ([5, 8192], [700, 17], [8192, 3]):
        yt = torch.tensor(y).unsqueeze(1)
        g, h, c = m.predictor_step(yt, h, c)
        with torch.no_grad():
            out, (h2, c2) = lstm(emb[yt[:, 0]].unsqueeze(0), (h2, c2))
        assert float((g[:, :, 0] - out[0]).abs().max()) < 1e-6
        assert float((h - h2).abs().max()) < 1e-6 and float((c - c2).abs().max()) < 1e-6


def test_first_streaming_chunk_equals_full_context(features_ref):
    """Chains the streaming step to the pinned full-context encoder: with an EMPTY cache and drop_extra_pre_encoded = 0 the
    cache-aware step attends over exactly the chunk's own tokens (the 256 cache positions are masked out, relative positions
    come from the longer 256 + Tq table through the same rel_shift), and its conv sees [zero cache(4) | chunk | 0000] -- the
    symmetric (4,4) padding of the full-context module.  So its emitted frames must equal offline() on the same frames."""
    m = ModelRef(model_dir(2))
    m.drop_pre = 0
    f = normalized_features(features_ref, 1.0, 77)[:, :57]
    f[0] = 0.0
    x = torch.from_numpy(f[None])
    cc, ct, cl = m.initial_cache(1)
    enc_s, el_s, cc1, ct1, cl1 = m.stream_step(x, torch.tensor([57]), cc, ct, cl)
    enc_o, el_o = m.offline(x, torch.tensor([57]))
    assert int(el_o) == 8 and enc_s.shape[2] == m.valid_out
    assert float((enc_s - enc_o[:, :, : m.valid_out]).abs().max()) < 1e-5
    # and the caches it hands on are what the next chunk needs: cache_len = tokens kept, time cache = last 4 kept GLU columns
    assert int(cl1) == 8 - m.drop


@pytest.mark.parametrize("cache_len", [0, 5, 97, 256])
def test_cache_aware_attention_step_matches_hf_attention(features_ref, cache_len):
    """Pins the part of the streaming step the full-context comparison cannot reach -- attention with Tq != Tk over
    [cache_last_channel || chunk], the 256 + Tq relative-position table, the validity mask of a partly filled cache and the FIFO cache
    update -- against Hugging Face's ParakeetEncoderAttention, an implementation that knows nothing about caches.

    Identity used: NeMo caches the attention INPUT (the norm_self_att output), so the chunk's attention output rows are the LAST Tq rows
    of plain full self-attention over the sequence [valid cache rows ; chunk rows] (scores depend on content and on the distance
    i - j only).  The oracle is made attention-only by zeroing the second linear of both FFNs and pointwise_conv2 (each layer becomes
    norm_out(x + MHA(norm_self_att(x)))); stream_step() then runs through its public contract with a random cache whose INVALID prefix
    holds garbage, and HF's attention module is driven layer by layer on the concatenated sequences."""
    m = ModelRef(model_dir(2))
    hf = _hf_encoder(m)
    for l in range(m.L):
        p = f"encoder.layers.{l}."
        for n in ("feed_forward1.linear2.weight", "feed_forward2.linear2.weight", "conv.pointwise_conv2.weight"):
            m.w[p + n] = torch.zeros_like(m.w[p + n])
    f = normalized_features(features_ref, 1.0, 31)[:, :57]
    f[0] = 0.0
    x_in = torch.from_numpy(f[None])
    torch.manual_seed(11 + cache_len)
    cc = torch.randn(1, m.L, m.S, m.D)                       # rows [S - cache_len, S) are the valid suffix; the rest must be ignored
    ct = torch.zeros(1, m.L, m.D, m.KT)
    enc, enc_len, cc_out, _, len_out = m.stream_step(x_in, torch.tensor([57]), cc, ct, torch.tensor([cache_len]))

    x, _ = m.pre_encode(x_in.transpose(1, 2), torch.tensor([57]))
    x = x[:, m.drop_pre:, :]
    Tq = x.size(1)
    keep = Tq - m.drop
    assert Tq == 6 and int(len_out) == min(cache_len + keep, m.S) and int(enc_len) == m.valid_out
    with torch.no_grad():
        for l in range(m.L):
            p = f"encoder.layers.{l}."
            a = torch.nn.functional.layer_norm(x, (m.D,), m.w[p + "norm_self_att.weight"], m.w[p + "norm_self_att.bias"], 1e-5)
            seq = torch.cat([cc[:, l, m.S - cache_len:, :], a], dim=1)                 # [1, cache_len + Tq, D]
            att, _ = hf.layers[l].self_attn(hidden_states=seq, position_embeddings=hf.encode_positions(seq), attention_mask=None)
            x = torch.nn.functional.layer_norm(x + att[:, -Tq:, :], (m.D,), m.w[p + "norm_out.weight"], m.w[p + "norm_out.bias"], 1e-5)
            # FIFO update: oldest `keep` rows leave, the chunk's first `keep` attention inputs enter at the END
            want_cache = torch.cat([cc[:, l, keep:, :], a[:, :keep, :]], dim=1)
            assert float((cc_out[:, l] - want_cache).abs().max()) < 1e-5
    d = (enc - x.transpose(1, 2)[:, :, : m.valid_out]).abs()
    assert float(d.max()) < 2e-5, float(d.max())
    assert float(enc.abs().mean()) > 0.1


def test_cached_convolution_step_matches_hf_convolution_module(features_ref):
    """The other streaming-specific piece: the conv module with its time cache.  NeMo's CausalConv1D with a cache convolves
    [cache_last_time(4) | chunk | 0000] without further padding and hands on new_x[..., :-cache_drop][..., -4:].  The cached columns are
    post-GLU activations of earlier frames, so for cached columns GLU(pointwise_conv1(h_prev)) the chunk's outputs are the LAST Tq rows of
    Hugging Face's ParakeetEncoderConvolutionModule (symmetric (4,4) zero padding, no notion of a cache) applied to the sequence
    [h_prev ; norm_conv(x)].  The oracle is made conv-only (attention linear_out and both FFN second linears zeroed: each layer is
    norm_out(x + Conv(norm_conv(x)))) and run through stream_step()'s contract."""
    m = ModelRef(model_dir(2))
    hf = _hf_encoder(m)
    for l in range(m.L):
        p = f"encoder.layers.{l}."
        for n in ("feed_forward1.linear2.weight", "feed_forward2.linear2.weight", "self_attn.linear_out.weight"):
            m.w[p + n] = torch.zeros_like(m.w[p + n])
    f = normalized_features(features_ref, 1.0, 32)[:, :57]
    f[0] = 0.0
    x_in = torch.from_numpy(f[None])
    torch.manual_seed(5)
    h_prev = torch.randn(m.L, 1, m.KT, m.D)                       # norm_conv outputs of the 4 frames before the chunk, per layer
    ct = torch.zeros(1, m.L, m.D, m.KT)
    for l in range(m.L):
        pw1 = m.w[f"encoder.layers.{l}.conv.pointwise_conv1.weight"]
        pw1 = pw1[:, :, 0] if pw1.dim() == 3 else pw1
        ct[0, l] = torch.nn.functional.glu(torch.nn.functional.linear(h_prev[l], pw1), dim=-1)[0].transpose(0, 1)
    cc = torch.zeros(1, m.L, m.S, m.D)
    enc, _, _, ct_out, _ = m.stream_step(x_in, torch.tensor([57]), cc, ct, torch.tensor([0]))

    x, _ = m.pre_encode(x_in.transpose(1, 2), torch.tensor([57]))
    x = x[:, m.drop_pre:, :]
    Tq = x.size(1)
    with torch.no_grad():
        for l in range(m.L):
            p = f"encoder.layers.{l}."
            c_in = torch.nn.functional.layer_norm(x, (m.D,), m.w[p + "norm_conv.weight"], m.w[p + "norm_conv.bias"], 1e-5)
            out = hf.layers[l].conv(torch.cat([h_prev[l], c_in], dim=1))[:, -Tq:, :]
            # time cache handed on: [x3, x4, x5, 0] of the chunk's post-GLU columns (docs/VALIDATION_REPORT_TRACE.md:212: last slot is zero)
            pw1 = m.w[p + "conv.pointwise_conv1.weight"]
            pw1 = pw1[:, :, 0] if pw1.dim() == 3 else pw1
            glu = torch.nn.functional.glu(torch.nn.functional.linear(c_in, pw1), dim=-1)[0].transpose(0, 1)      # [D, Tq]
            want_ct = torch.cat([glu, torch.zeros(m.D, 4 - m.drop)], dim=1)[:, -m.KT:]      # [..|chunk|0000] minus its last cache_drop columns
            assert want_ct.shape == (m.D, m.KT) and float((ct_out[0, l] - want_ct).abs().max()) < 1e-5
            assert float(ct_out[0, l, :, -1].abs().max()) == 0.0
            x = torch.nn.functional.layer_norm(x + out, (m.D,), m.w[p + "norm_out.weight"], m.w[p + "norm_out.bias"], 1e-5)
    d = (enc - x.transpose(1, 2)[:, :, : m.valid_out]).abs()
    assert float(d.max()) < 2e-5, float(d.max())
    assert float(enc.abs().mean()) > 0.1
