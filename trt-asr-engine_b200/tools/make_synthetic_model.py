#!/usr/bin/env python3
"""Write a <model_dir> (weights.bin + vocab.txt + model_meta.json) with seeded random weights of the
published Parakeet-TDT-0.6B-v3 architecture.

The real checkpoint (.nemo, sha256 3cbdc858..., /root/reference/contracts/parakeet-tdt-0.6b-v3.contract.json:6)
is not available offline, so BASELINE.json asks for "random-init weights of the published architecture".
Tensor names/shapes follow the NeMo state_dict (architecture constants from
/root/reference/audit_model_arch.json:12-47 and the contract file :54-66, :161-215).

Init is chosen so that activations stay O(1) through all layers and the greedy TDT decode behaves like
speech (mostly blank, occasional tokens, durations mostly 1-2) instead of hitting the 8-symbol cap:
  * linear / pointwise weights ~ N(0, 1/fan_in); LayerNorm gamma 1+0.1N, beta 0.1N
  * BatchNorm running stats random but benign; pos_bias_u/v ~ 0.3N (NeMo zero-inits them; random
    values exercise the path)
  * joint output rows get log-normal row scales (a few "confident" tokens) and the blank / duration
    biases are shifted (values recorded in model_meta.json)
Every GEMM weight is stored as bf16, so the CPU oracle and the GPU path read identical numbers.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weights_io import DT_BF16, DT_F32, write_weights  # noqa: E402

D_MODEL, N_HEADS, FF_DIM, CONV_K, SUB_CH, FEAT_IN = 1024, 8, 4096, 9, 256, 128
VOCAB, N_DUR, PRED_H, PRED_L, JOINT_H = 8193, 5, 640, 2, 640
BLANK = 8192

SPECIALS = ["<unk>", "<|nospeech|>", "<pad>", "<|endoftext|>", "<|startoftranscript|>", "<|en|>",
            "<|nopnc|>", "<|noitn|>"]


def synth_vocab(rng: np.random.Generator) -> list[str]:
    """8192 SentencePiece-style pieces: specials first, a few punctuation-only pieces, then syllables.
    Line index == token id (reference: cpp/src/tokenizer.cpp:9-23)."""
    pieces = list(SPECIALS) + [".", ",", "?", "▁-", "'"]
    cons, vow = "bcdfghjklmnprstvwz", "aeiou"
    seen = set(pieces)
    while len(pieces) < VOCAB - 1:
        n = int(rng.integers(1, 4))
        s = "".join(cons[int(rng.integers(len(cons)))] + vow[int(rng.integers(len(vow)))] for _ in range(n))
        if rng.random() < 0.45:
            s = "▁" + s
        if s not in seen:
            seen.add(s)
            pieces.append(s)
    return pieces  # blank (8192) has no line, like the real vocab.txt (8192 lines)


def build(out_dir: str, n_layers: int, seed: int, blank_rate: float, row_sigma: float, logit_gain: float):
    os.makedirs(out_dir, exist_ok=True)
    g = torch.Generator().manual_seed(seed)
    T = {}

    def randn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g) * std).numpy()

    def add(name, arr, dt=DT_F32):
        T[name] = (np.asarray(arr, dtype=np.float32), dt)

    # --- pre_encode (dw_striding x8, 256 channels) ---
    add("encoder.pre_encode.conv.0.weight", randn(SUB_CH, 1, 3, 3, std=1 / 3))
    add("encoder.pre_encode.conv.0.bias", randn(SUB_CH, std=0.1))
    for dw, pw in ((2, 3), (5, 6)):
        add(f"encoder.pre_encode.conv.{dw}.weight", randn(SUB_CH, 1, 3, 3, std=1 / 3))
        add(f"encoder.pre_encode.conv.{dw}.bias", randn(SUB_CH, std=0.1))
        add(f"encoder.pre_encode.conv.{pw}.weight", randn(SUB_CH, SUB_CH, 1, 1, std=(2.0 / SUB_CH) ** 0.5), DT_BF16)
        add(f"encoder.pre_encode.conv.{pw}.bias", randn(SUB_CH, std=0.1))
    add("encoder.pre_encode.out.weight", randn(D_MODEL, SUB_CH * 16, std=2.0 / (SUB_CH * 16) ** 0.5), DT_BF16)
    add("encoder.pre_encode.out.bias", randn(D_MODEL, std=0.1))

    # --- conformer layers ---
    for i in range(n_layers):
        p = f"encoder.layers.{i}."
        for nm in ("norm_feed_forward1", "norm_self_att", "norm_conv", "norm_feed_forward2", "norm_out"):
            add(p + nm + ".weight", 1.0 + randn(D_MODEL, std=0.1))
            add(p + nm + ".bias", randn(D_MODEL, std=0.1))
        for ff in ("feed_forward1", "feed_forward2"):
            add(p + ff + ".linear1.weight", randn(FF_DIM, D_MODEL, std=D_MODEL ** -0.5), DT_BF16)
            add(p + ff + ".linear2.weight", randn(D_MODEL, FF_DIM, std=FF_DIM ** -0.5), DT_BF16)
        for nm in ("linear_q", "linear_k", "linear_v", "linear_out", "linear_pos"):
            add(p + f"self_attn.{nm}.weight", randn(D_MODEL, D_MODEL, std=D_MODEL ** -0.5), DT_BF16)
        add(p + "self_attn.pos_bias_u", randn(N_HEADS, D_MODEL // N_HEADS, std=0.3))
        add(p + "self_attn.pos_bias_v", randn(N_HEADS, D_MODEL // N_HEADS, std=0.3))
        add(p + "conv.pointwise_conv1.weight", randn(2 * D_MODEL, D_MODEL, 1, std=D_MODEL ** -0.5), DT_BF16)
        add(p + "conv.depthwise_conv.weight", randn(D_MODEL, 1, CONV_K, std=1 / 3))
        add(p + "conv.batch_norm.weight", 1.0 + randn(D_MODEL, std=0.1))
        add(p + "conv.batch_norm.bias", randn(D_MODEL, std=0.1))
        add(p + "conv.batch_norm.running_mean", randn(D_MODEL, std=0.1))
        add(p + "conv.batch_norm.running_var", 0.5 + torch.rand(D_MODEL, generator=g).numpy())
        add(p + "conv.pointwise_conv2.weight", randn(D_MODEL, D_MODEL, 1, std=D_MODEL ** -0.5), DT_BF16)

    # --- predictor ---
    emb = randn(VOCAB, PRED_H, std=1.0)
    emb[BLANK] = 0.0  # padding_idx = blank (blank_as_pad)
    add("decoder.prediction.embed.weight", emb, DT_BF16)
    for l in range(PRED_L):
        add(f"decoder.prediction.dec_rnn.lstm.weight_ih_l{l}", randn(4 * PRED_H, PRED_H, std=PRED_H ** -0.5), DT_BF16)
        add(f"decoder.prediction.dec_rnn.lstm.weight_hh_l{l}", randn(4 * PRED_H, PRED_H, std=PRED_H ** -0.5), DT_BF16)
        add(f"decoder.prediction.dec_rnn.lstm.bias_ih_l{l}", randn(4 * PRED_H, std=0.1))
        add(f"decoder.prediction.dec_rnn.lstm.bias_hh_l{l}", randn(4 * PRED_H, std=0.1))

    # --- joint ---
    add("joint.enc.weight", randn(JOINT_H, D_MODEL, std=D_MODEL ** -0.5), DT_BF16)
    add("joint.enc.bias", randn(JOINT_H, std=0.1))
    add("joint.pred.weight", randn(JOINT_H, PRED_H, std=2.0 * PRED_H ** -0.5), DT_BF16)
    add("joint.pred.bias", randn(JOINT_H, std=0.1))
    w_out = randn(VOCAB + N_DUR, JOINT_H, std=logit_gain * JOINT_H ** -0.5)
    row_scale = np.exp(row_sigma * randn(VOCAB + N_DUR))
    row_scale[BLANK:] = 1.0
    w_out *= row_scale[:, None]
    b_out = randn(VOCAB + N_DUR, std=0.1)
    # Calibrate the blank / duration biases on a proxy of the joint's hidden state, relu(enc_proj(x) + pred_proj(g)) with
    # x ~ LayerNorm output and g ~ LSTM output, so that greedy decode behaves like speech: blank wins `blank_rate` of the
    # steps and durations are mostly 1-2 (otherwise random logits emit a token on every step and hit the 8-symbol cap).
    xs = randn(768, D_MODEL)
    gs = np.tanh(randn(768, PRED_H)) * 0.5
    hid = np.maximum(xs @ T["joint.enc.weight"][0].T + T["joint.enc.bias"][0] + gs @ T["joint.pred.weight"][0].T
                     + T["joint.pred.bias"][0], 0.0).astype(np.float32)
    tok_max = (hid @ w_out[:BLANK].T + b_out[:BLANK]).max(axis=1)
    blank_raw = hid @ w_out[BLANK] + b_out[BLANK]
    b_out[BLANK] += float(np.quantile(tok_max - blank_raw, blank_rate))
    dur_sd = float((hid @ w_out[VOCAB:].T).std())
    b_out[VOCAB:] += np.array([-1.0, 1.5, 1.2, 0.0, -1.0], dtype=np.float32) * dur_sd
    add("joint.joint_net.2.weight", w_out, DT_BF16)
    add("joint.joint_net.2.bias", b_out)

    cfg = dict(n_layers=n_layers, d_model=D_MODEL, n_heads=N_HEADS, ff_dim=FF_DIM, conv_kernel=CONV_K,
               sub_channels=SUB_CH, feat_in=FEAT_IN, vocab=VOCAB, n_dur=N_DUR, pred_hidden=PRED_H,
               pred_layers=PRED_L, joint_hidden=JOINT_H, blank_id=BLANK, cache_size=256, time_ctx=4,
               cache_drop=3, valid_out_len=3, drop_extra_pre_encoded=2, max_symbols=8, seed=seed)
    write_weights(os.path.join(out_dir, "weights.bin"), cfg, T)

    vocab = synth_vocab(np.random.default_rng(seed + 17))
    with open(os.path.join(out_dir, "vocab.txt"), "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    meta = dict(cfg, generator="make_synthetic_model.py", gen_version=GEN_VERSION, blank_rate=blank_rate, row_sigma=row_sigma,
                logit_gain=logit_gain, duration_values=[0, 1, 2, 3, 4],
                note="seeded random weights; GEMM weights bf16-representable")
    with open(os.path.join(out_dir, "model_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    n_params = sum(int(np.prod(a.shape)) for a, _ in T.values())
    return n_params


def ensure_model(out_dir: str, n_layers: int = 24, seed: int = 0, **kw) -> str:
    """Create the model dir if missing (used by tests / bench on a fresh box)."""
    meta = os.path.join(out_dir, "model_meta.json")
    if os.path.exists(meta) and os.path.exists(os.path.join(out_dir, "weights.bin")):
        with open(meta) as f:
            m = json.load(f)
        if m.get("n_layers") == n_layers and m.get("seed") == seed and m.get("gen_version") == GEN_VERSION:
            return out_dir
    kw.setdefault("blank_rate", DEFAULTS["blank_rate"])
    kw.setdefault("row_sigma", DEFAULTS["row_sigma"])
    kw.setdefault("logit_gain", DEFAULTS["logit_gain"])
    build(out_dir, n_layers, seed, **kw)
    return out_dir


DEFAULTS = dict(blank_rate=0.75, row_sigma=0.6, logit_gain=3.0)
GEN_VERSION = 2

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--blank-rate", type=float, default=DEFAULTS["blank_rate"])
    ap.add_argument("--row-sigma", type=float, default=DEFAULTS["row_sigma"])
    ap.add_argument("--logit-gain", type=float, default=DEFAULTS["logit_gain"])
    a = ap.parse_args()
    n = build(a.out, a.layers, a.seed, a.blank_rate, a.row_sigma, a.logit_gain)
    print(f"wrote {a.out}: {n} parameters, {a.layers} layers")
