#!/usr/bin/env python3
"""convert_nemo.py -- .nemo checkpoint -> <model_dir>/{weights.bin, vocab.txt, model_meta.json} for libparakeet_trt (B200 build).

Takes the place of the reference's NeMo -> ONNX -> TensorRT chain (/root/reference/tools/export_onnx/export.py:854-904 for
the tokenizer assets, :970-997 for the metadata, tools/build_trt/ for the engines): this build needs no graph export, only
the tensors.  A .nemo file is a tar archive (optionally gzip) holding `model_config.yaml`, `model_weights.ckpt` (a PyTorch
state_dict) and the tokenizer files; weights.bin keeps the NeMo state_dict names (weights_io.py),
so the conversion is a filtered copy: GEMM operands to bf16, norms / biases / depthwise kernels f32.

NeMo itself is not needed.  The published checkpoint (parakeet-tdt-0.6b-v3.nemo, sha256 3cbdc858...,
contracts/parakeet-tdt-0.6b-v3.contract.json:6) is not available offline; tests/test_convert_nemo.py round-trips a
synthetic archive of the same structure.
"""
from __future__ import annotations

import argparse
import io
import json
import os
import re
import sys
import tarfile
from typing import Dict, List, Optional, Tuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weights_io import DT_BF16, DT_F32, write_weights  # noqa: E402

# tensors the runtime reads (regexes over NeMo state_dict keys) and whether they are GEMM operands (bf16)
_WANTED: List[Tuple[str, int]] = [
    (r"encoder\.pre_encode\.conv\.(0|2|5)\.(weight|bias)", DT_F32),
    (r"encoder\.pre_encode\.conv\.(3|6)\.weight", DT_BF16),
    (r"encoder\.pre_encode\.conv\.(3|6)\.bias", DT_F32),
    (r"encoder\.pre_encode\.out\.weight", DT_BF16),
    (r"encoder\.pre_encode\.out\.bias", DT_F32),
    (r"encoder\.layers\.\d+\.norm_(feed_forward1|self_att|conv|feed_forward2|out)\.(weight|bias)", DT_F32),
    (r"encoder\.layers\.\d+\.feed_forward[12]\.linear[12]\.weight", DT_BF16),
    (r"encoder\.layers\.\d+\.self_attn\.linear_(q|k|v|out|pos)\.weight", DT_BF16),
    (r"encoder\.layers\.\d+\.self_attn\.pos_bias_[uv]", DT_F32),
    (r"encoder\.layers\.\d+\.conv\.pointwise_conv[12]\.weight", DT_BF16),
    (r"encoder\.layers\.\d+\.conv\.depthwise_conv\.weight", DT_F32),
    (r"encoder\.layers\.\d+\.conv\.batch_norm\.(weight|bias|running_mean|running_var)", DT_F32),
    (r"decoder\.prediction\.embed\.weight", DT_BF16),
    (r"decoder\.prediction\.dec_rnn\.lstm\.weight_(ih|hh)_l\d+", DT_BF16),
    (r"decoder\.prediction\.dec_rnn\.lstm\.bias_(ih|hh)_l\d+", DT_F32),
    (r"joint\.(enc|pred)\.weight", DT_BF16),
    (r"joint\.(enc|pred)\.bias", DT_F32),
    (r"joint\.joint_net\.2\.weight", DT_BF16),
    (r"joint\.joint_net\.2\.bias", DT_F32),
]
_WANTED_RE = [(re.compile("^" + p + "$"), dt) for p, dt in _WANTED]


def _member(tar: tarfile.TarFile, suffix: str) -> Optional[tarfile.TarInfo]:
    for m in tar.getmembers():
        if m.isfile() and os.path.basename(m.name).endswith(suffix):
            return m
    return None


def pieces_from_spm(model_bytes: bytes) -> List[str]:
    """SentencePiece model -> pieces in id order (what NeMo exposes as decoder.vocabulary for SPE tokenizers)."""
    import sentencepiece as spm
    sp = spm.SentencePieceProcessor()
    sp.LoadFromSerializedProto(model_bytes)
    return [sp.IdToPiece(i) for i in range(sp.GetPieceSize())]


def _cfg_get(d: dict, path: str, default=None):
    cur = d
    for k in path.split("."):
        if not isinstance(cur, dict) or k not in cur:
            return default
        cur = cur[k]
    return cur


def convert(nemo_path: str, out_dir: str, allow_vocab_mismatch: bool = False) -> Dict[str, int]:
    import torch
    import yaml
    os.makedirs(out_dir, exist_ok=True)
    with tarfile.open(nemo_path, "r:*") as tar:
        m_cfg, m_ckpt = _member(tar, "model_config.yaml"), _member(tar, "model_weights.ckpt")
        if m_cfg is None or m_ckpt is None:
            raise ValueError(f"{nemo_path}: not a .nemo archive (model_config.yaml / model_weights.ckpt missing)")
        ycfg = yaml.safe_load(tar.extractfile(m_cfg).read()) or {}
        sd = torch.load(io.BytesIO(tar.extractfile(m_ckpt).read()), map_location="cpu", weights_only=True)
        m_spm, m_vocab = _member(tar, "tokenizer.model"), _member(tar, "vocab.txt")
        if m_spm is not None:
            vocab = pieces_from_spm(tar.extractfile(m_spm).read())
        elif m_vocab is not None:
            vocab = tar.extractfile(m_vocab).read().decode("utf-8").splitlines()
        else:
            labels = _cfg_get(ycfg, "joint.vocabulary") or _cfg_get(ycfg, "labels") or []
            vocab = [str(x) for x in labels]
    if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]

    tensors: Dict[str, Tuple[np.ndarray, int]] = {}
    for name, t in sd.items():
        for rx, dt in _WANTED_RE:
            if rx.match(name):
                tensors[name] = (t.detach().to(torch.float32).cpu().numpy(), dt)
                break
    n_layers = 1 + max((int(re.match(r"encoder\.layers\.(\d+)\.", n).group(1)) for n in tensors if n.startswith("encoder.layers.")),
                       default=-1)
    need = ["encoder.pre_encode.out.weight", "decoder.prediction.embed.weight", "joint.joint_net.2.weight", "joint.enc.weight"]
    missing = [n for n in need if n not in tensors]
    if n_layers <= 0 or missing:
        raise ValueError(f"checkpoint lacks expected tensors: layers={n_layers} missing={missing}")
    per_layer = 5 * 2 + 4 + 5 + 2 + 2 + 1 + 4      # norms, FFN, attention, pos biases, pointwise, depthwise, batch-norm
    have = sum(1 for n in tensors if n.startswith("encoder.layers.0."))
    if have != per_layer:
        raise ValueError(f"layer 0 has {have} of {per_layer} expected tensors (use_bias / norm type differ from Parakeet-TDT-0.6B-v3?)")

    d_model = tensors["encoder.pre_encode.out.weight"][0].shape[0]
    sub_ch = tensors["encoder.pre_encode.conv.0.weight"][0].shape[0]
    vocab_p1, pred_h = tensors["decoder.prediction.embed.weight"][0].shape          # vocabulary + blank
    joint_out, joint_h = tensors["joint.joint_net.2.weight"][0].shape
    n_dur = joint_out - vocab_p1
    pred_l = 1 + max(int(re.search(r"_l(\d+)$", n).group(1)) for n in tensors if "dec_rnn.lstm.weight_ih" in n)
    n_heads = tensors["encoder.layers.0.self_attn.pos_bias_u"][0].shape[0]
    cfg = dict(n_layers=n_layers, d_model=d_model, n_heads=n_heads,
               ff_dim=tensors["encoder.layers.0.feed_forward1.linear1.weight"][0].shape[0],
               conv_kernel=tensors["encoder.layers.0.conv.depthwise_conv.weight"][0].shape[-1], sub_channels=sub_ch,
               feat_in=tensors["encoder.pre_encode.out.weight"][0].shape[1] // sub_ch * 8, vocab=vocab_p1, n_dur=n_dur,
               pred_hidden=pred_h, pred_layers=pred_l, joint_hidden=joint_h, blank_id=vocab_p1 - 1,
               # cache-aware streaming parameters of the reference's export (contract.json:255-266, export.py:98-106, 678)
               cache_size=256, time_ctx=4, cache_drop=3, valid_out_len=3, drop_extra_pre_encoded=2, max_symbols=8, seed=-1)
    ycfg_layers = _cfg_get(ycfg, "encoder.n_layers")
    if ycfg_layers is not None and int(ycfg_layers) != n_layers:
        raise ValueError(f"model_config.yaml says {ycfg_layers} layers, checkpoint holds {n_layers}")
    if len(vocab) != vocab_p1 - 1 and not allow_vocab_mismatch:
        raise ValueError(f"tokenizer has {len(vocab)} pieces, the embedding expects {vocab_p1 - 1} (+ blank)")

    write_weights(os.path.join(out_dir, "weights.bin"), cfg, tensors)
    with open(os.path.join(out_dir, "vocab.txt"), "w", encoding="utf-8") as f:      # one piece per line, line index == token id
        f.write("\n".join(vocab) + "\n")
    durations = _cfg_get(ycfg, "model_defaults.tdt_durations") or _cfg_get(ycfg, "loss.tdt_kwargs.durations") or list(range(n_dur))
    meta = dict(cfg, generator="convert_nemo.py", source=os.path.basename(nemo_path), duration_values=[int(x) for x in durations],
                n_tensors=len(tensors), note="GEMM operands stored as bf16 (round-to-nearest-even)")
    with open(os.path.join(out_dir, "model_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return cfg


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model", required=True, help="path to the .nemo file")
    ap.add_argument("--out", required=True, help="model directory to create")
    ap.add_argument("--allow-vocab-mismatch", action="store_true")
    a = ap.parse_args()
    c = convert(a.model, a.out, a.allow_vocab_mismatch)
    print(f"wrote {a.out}: {c['n_layers']} layers, d_model {c['d_model']}, vocab {c['vocab']} (+{c['n_dur']} durations)")
