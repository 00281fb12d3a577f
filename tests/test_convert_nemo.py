"""tools/convert_nemo.py on a synthetic .nemo archive of the published structure (tar: model_config.yaml + model_weights.ckpt +
tokenizer files): the conversion must reproduce the model directory bit for bit."""
import io
import os
import tarfile

import numpy as np
import pytest
import torch

from convert_nemo import convert, pieces_from_spm
from weights_io import read_weights


def _add(tar, name, data: bytes):
    info = tarfile.TarInfo(name)
    info.size = len(data)
    tar.addfile(info, io.BytesIO(data))


def _tiny_state_dict(seed=0):
    """NeMo state_dict keys of a Parakeet-TDT-shaped model with tiny dimensions (values are multiples of 1/8: bf16-exact)."""
    g = torch.Generator().manual_seed(seed)
    D, H, FF, K, C, V, ND, PH, JH = 16, 2, 32, 9, 4, 12, 5, 8, 8

    def r(*shape):
        return torch.randint(-16, 17, shape, generator=g).to(torch.float32) / 8.0

    sd = {}
    for i in (0, 2, 5):
        sd[f"encoder.pre_encode.conv.{i}.weight"], sd[f"encoder.pre_encode.conv.{i}.bias"] = r(C, 1, 3, 3), r(C)
    for i in (3, 6):
        sd[f"encoder.pre_encode.conv.{i}.weight"], sd[f"encoder.pre_encode.conv.{i}.bias"] = r(C, C, 1, 1), r(C)
    sd["encoder.pre_encode.out.weight"], sd["encoder.pre_encode.out.bias"] = r(D, C * 16), r(D)
    for l in range(2):
        p = f"encoder.layers.{l}."
        for n in ("norm_feed_forward1", "norm_self_att", "norm_conv", "norm_feed_forward2", "norm_out"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = r(D), r(D)
        for ff in ("feed_forward1", "feed_forward2"):
            sd[p + ff + ".linear1.weight"], sd[p + ff + ".linear2.weight"] = r(FF, D), r(D, FF)
        for n in ("q", "k", "v", "out", "pos"):
            sd[p + f"self_attn.linear_{n}.weight"] = r(D, D)
        sd[p + "self_attn.pos_bias_u"], sd[p + "self_attn.pos_bias_v"] = r(H, D // H), r(H, D // H)
        sd[p + "conv.pointwise_conv1.weight"], sd[p + "conv.pointwise_conv2.weight"] = r(2 * D, D, 1), r(D, D, 1)
        sd[p + "conv.depthwise_conv.weight"] = r(D, 1, K)
        for n in ("weight", "bias", "running_mean", "running_var"):
            sd[p + "conv.batch_norm." + n] = r(D)
        sd[p + "conv.batch_norm.num_batches_tracked"] = torch.tensor(7)          # junk the runtime does not read
    sd["decoder.prediction.embed.weight"] = r(V + 1, PH)
    for l in range(2):
        for n in ("ih", "hh"):
            sd[f"decoder.prediction.dec_rnn.lstm.weight_{n}_l{l}"], sd[f"decoder.prediction.dec_rnn.lstm.bias_{n}_l{l}"] = r(4 * PH, PH), r(4 * PH)
    sd["joint.enc.weight"], sd["joint.enc.bias"] = r(JH, D), r(JH)
    sd["joint.pred.weight"], sd["joint.pred.bias"] = r(JH, PH), r(JH)
    sd["joint.joint_net.2.weight"], sd["joint.joint_net.2.bias"] = r(V + 1 + ND, JH), r(V + 1 + ND)
    sd["preprocessor.featurizer.window"] = torch.zeros(400)                       # more junk
    sd["encoder.pos_enc.pe"] = torch.zeros(1, 99, D)
    return sd, dict(n_layers=2, d_model=D, n_heads=H, ff_dim=FF, conv_kernel=K, sub_channels=C, feat_in=128, vocab=V + 1, n_dur=ND,
                    pred_hidden=PH, pred_layers=2, joint_hidden=JH, blank_id=V)


def _fake_nemo(path, vocab_lines, vocab_name="a1b2_vocab.txt", gz=True):
    sd, cfg = _tiny_state_dict()
    buf = io.BytesIO()
    torch.save(sd, buf)
    yaml_txt = f"encoder:\n  n_layers: {cfg['n_layers']}\n  d_model: {cfg['d_model']}\nmodel_defaults:\n  tdt_durations: [0, 1, 2, 3, 4]\n"
    with tarfile.open(path, "w:gz" if gz else "w") as tar:
        _add(tar, "./model_config.yaml", yaml_txt.encode())
        _add(tar, "./model_weights.ckpt", buf.getvalue())
        _add(tar, "./" + vocab_name, ("\n".join(vocab_lines) + "\n").encode())
    return sd, cfg


VOCAB12 = ["<unk>", "<pad>", "<|startoftranscript|>", "<|en|>", ".", "▁a", "▁b", "c", "d", "▁e", "f", "g"]


def test_round_trip(tmp_path):
    sd, cfg = _fake_nemo(str(tmp_path / "m.nemo"), VOCAB12)
    out = str(tmp_path / "converted")
    got_cfg = convert(str(tmp_path / "m.nemo"), out)
    cfg2, w2 = read_weights(os.path.join(out, "weights.bin"))
    for k, v in cfg.items():
        assert cfg2[k] == v == got_cfg[k], k
    assert (cfg2["cache_size"], cfg2["time_ctx"], cfg2["cache_drop"], cfg2["valid_out_len"], cfg2["drop_extra_pre_encoded"]) == (256, 4, 3, 3, 2)
    wanted = {k for k in sd if not ("num_batches_tracked" in k or k.startswith("preprocessor.") or k.startswith("encoder.pos_enc."))}
    assert set(w2) == wanted                                   # junk keys dropped, nothing else lost
    for k in wanted:
        assert w2[k].shape == tuple(sd[k].shape) and np.array_equal(w2[k], sd[k].numpy()), k
    assert open(os.path.join(out, "vocab.txt"), encoding="utf-8").read().splitlines() == VOCAB12


def test_rejects_foreign_archives(tmp_path):
    with tarfile.open(str(tmp_path / "x.nemo"), "w") as tar:
        _add(tar, "readme.txt", b"hello")
    with pytest.raises(ValueError, match="not a .nemo archive"):
        convert(str(tmp_path / "x.nemo"), str(tmp_path / "o"))
    # tokenizer / embedding size mismatch is an error unless explicitly allowed
    _fake_nemo(str(tmp_path / "m2.nemo"), ["a", "b", "c"], gz=False)
    with pytest.raises(ValueError, match="pieces"):
        convert(str(tmp_path / "m2.nemo"), str(tmp_path / "o2"))
    convert(str(tmp_path / "m2.nemo"), str(tmp_path / "o3"), allow_vocab_mismatch=True)


def test_sentencepiece_vocab(tmp_path):
    import sentencepiece as spm
    corpus = tmp_path / "c.txt"
    corpus.write_text("\n".join("the quick brown fox jumps over the lazy dog number %d" % i for i in range(200)))
    spm.SentencePieceTrainer.Train(input=str(corpus), model_prefix=str(tmp_path / "tok"), vocab_size=60, model_type="bpe",
                                   minloglevel=2)
    pieces = pieces_from_spm(open(str(tmp_path / "tok.model"), "rb").read())
    assert len(pieces) == 60 and pieces[0] == "<unk>" and any(p.startswith("▁") for p in pieces)
