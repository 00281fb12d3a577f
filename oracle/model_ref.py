"""oracle/model_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (PyTorch fp32) restatement of the model arithmetic behind the reference's three TensorRT engines.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

The arithmetic itself lives in an un-vendored dependency, nemo_toolkit[asr]==2.6.0
(/root/reference/tools/export_onnx/requirements.txt:2, /root/reference/audit_model_arch.json:8), reached by the
reference only through these call sites, which this file follows:
  * encoder.forward_for_export(audio_signal, length, cache_last_channel, cache_last_time, cache_last_channel_len)
        /root/reference/tools/export_onnx/export.py:343-375        -> EncoderRef.stream_step
  * encoder(audio_signal, length)  (offline)  tools/verify_nemo/tdt_trace.py:272-274 -> EncoderRef.offline
  * decoder.predict(y, state=[h,c], add_sos=False)  export.py:580-599            -> PredictorRef.step
  * joint(encoder_outputs=, decoder_outputs=) with log_softmax off  export.py:601-612 -> JointRef.logits
  * greedy TDT loop  tools/verify_nemo/tdt_trace.py:277-353 == cpp/src/parakeet_trt.cpp:2914-3676 -> tdt_greedy_chunk
  * 41/57-frame schedule  tools/verify_nemo/streaming_encoder_reference.py:522-550 -> streaming_schedule

PARITY UNPINNED AGAINST THE REFERENCE'S OWN OUTPUTS: NeMo, the .nemo weights and the golden JSONL tensors
(artifacts/reference/, git-ignored) are all absent, so this restates NeMo 2.6.0's published module semantics
(SURVEY.md section 8a [UPSTREAM]).  What pins it instead:
  * numerically, against independent public implementations that ARE in this image (tests/test_oracle_vs_hf.py): the
    full-context encoder equals transformers 5.5 `ParakeetEncoder` (Hugging Face's port of the NeMo FastConformer encoder)
    on the same seeded weights to 2e-6; the predictor equals torch.nn.LSTM (what NeMo's RNNTDecoder wraps) to 1e-6.
    The cache-aware streaming step runs the same `_layer()` as the full-context encoder, and with an empty cache and no
    dropped tokens its emitted frames equal offline() on the same frames (same test file);
  * the greedy TDT loop: the reference's own tools/verify_nemo/tdt_trace.py, executed unmodified on these modules, produced
    tests/golden/tdt_trace_ref.json; tdt_greedy_chunk + prime reproduce it step for step (tests/test_oracle_kats.py);
  * structurally, by checked-in reference evidence (tests/test_oracle_kats.py): layouts, the schedule, encoded_lengths=3,
    cache_len sequences 1,4,7,... and 1,2,3,4 (docs/VALIDATION_REPORT_TRACE.md:173-177, 209-213), conv-cache last column
    zero (:212).
"""
from __future__ import annotations

import math
import os
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "trt-asr-engine_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from weights_io import read_weights  # noqa: E402  (file-format reader only; no product compute)

INF_VAL = 10000.0  # NeMo multi_head_attention.INF_VAL


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class ModelRef:
    """Loads weights.bin and exposes encoder / predictor / joint restatements.

    act_round: None -> pure fp32 (the oracle proper).  "bf16" -> GEMM input activations are rounded to bf16
    first (a CPU *simulation* of the tensor-core path's input rounding, used only to choose tolerances).
    """

    def __init__(self, model_dir: str, act_round: Optional[str] = None):
        cfg, w = read_weights(os.path.join(model_dir, "weights.bin"))
        self.cfg = cfg
        self.w: Dict[str, torch.Tensor] = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}
        self.L = cfg["n_layers"]
        self.D = cfg["d_model"]
        self.H = cfg["n_heads"]
        self.dk = self.D // self.H
        self.S = cfg["cache_size"]
        self.KT = cfg["time_ctx"]
        self.drop = cfg["cache_drop"]
        self.valid_out = cfg["valid_out_len"]
        self.drop_pre = cfg["drop_extra_pre_encoded"]
        self.blank = cfg["blank_id"]
        self.vocab = cfg["vocab"]
        self.n_dur = cfg["n_dur"]
        self.act_round = act_round
        self._pe_cache: Dict[int, torch.Tensor] = {}
        with open(os.path.join(model_dir, "vocab.txt"), encoding="utf-8") as f:
            self.vocab_lines = [ln.rstrip("\n").rstrip("\r") for ln in f]

    # ------------------------------------------------------------------ helpers
    def _lin(self, x, wname, bname=None):
        if self.act_round == "bf16":
            x = bf16_round(x)
        wt = self.w[wname]
        if wt.dim() == 3:       # Conv1d kernel_size 1 == Linear over channels
            wt = wt[:, :, 0]
        return F.linear(x, wt, self.w[bname] if bname else None)

    @staticmethod
    def calc_length(L: torch.Tensor, repeat: int = 3) -> torch.Tensor:
        # NeMo conv_subsampling.calc_length: floor((L + 2 - 3)/2 + 1), three times
        for _ in range(repeat):
            L = torch.div(L.to(torch.float32) + 2 - 3, 2).add(1.0).floor()
        return L.to(torch.int64)

    def rel_pos_emb(self, length: int) -> torch.Tensor:
        """RelPositionalEncoding: rows for relative positions length-1 ... -(length-1); sin even / cos odd."""
        if length not in self._pe_cache:
            pos = torch.arange(length - 1, -length, -1, dtype=torch.float32).unsqueeze(1)
            div = torch.exp(torch.arange(0, self.D, 2, dtype=torch.float32) * -(math.log(INF_VAL) / self.D))
            pe = torch.zeros(pos.size(0), self.D)
            pe[:, 0::2] = torch.sin(pos * div)
            pe[:, 1::2] = torch.cos(pos * div)
            self._pe_cache[length] = pe.unsqueeze(0)
        return self._pe_cache[length]

    # ------------------------------------------------------------------ pre-encode (ConvSubsampling dw_striding)
    def pre_encode(self, x_btf: torch.Tensor, lengths: torch.Tensor):
        w = self.w
        p = "encoder.pre_encode."
        x = x_btf.unsqueeze(1)
        x = F.relu(F.conv2d(x, w[p + "conv.0.weight"], w[p + "conv.0.bias"], stride=2, padding=1))
        for dw, pw in ((2, 3), (5, 6)):
            x = F.conv2d(x, w[p + f"conv.{dw}.weight"], w[p + f"conv.{dw}.bias"], stride=2, padding=1, groups=x.size(1))
            if self.act_round == "bf16":
                x = bf16_round(x)
            x = F.relu(F.conv2d(x, w[p + f"conv.{pw}.weight"], w[p + f"conv.{pw}.bias"]))
        b, c, t, f = x.shape
        x = self._lin(x.transpose(1, 2).reshape(b, t, c * f), p + "out.weight", p + "out.bias")
        return x, self.calc_length(lengths)

    # ------------------------------------------------------------------ masks (ConformerEncoder._create_masks, regular, [-1,-1])
    @staticmethod
    def _create_masks(padding_length, max_len, offset):
        ar = torch.arange(0, max_len).expand(padding_length.size(0), -1)
        pad = ar < padding_length.unsqueeze(-1)
        if offset is not None:
            pad = (ar >= offset.unsqueeze(-1)) & pad
        pm = pad.unsqueeze(1).repeat(1, max_len, 1)
        pm = pm & pm.transpose(1, 2)
        att_mask = ~pm              # True = masked
        return ~pad, att_mask

    # ------------------------------------------------------------------ one ConformerLayer
    def _layer(self, i, x, att_mask, pos_emb, pad_mask, cache_ch, cache_tm):
        w = self.w
        p = f"encoder.layers.{i}."
        D, H, dk = self.D, self.H, self.dk

        def ln(t, nm):
            return F.layer_norm(t, (D,), w[p + nm + ".weight"], w[p + nm + ".bias"], 1e-5)

        def ff(t, nm):
            return self._lin(F.silu(self._lin(t, p + nm + ".linear1.weight")), p + nm + ".linear2.weight")

        residual = x + 0.5 * ff(ln(x, "norm_feed_forward1"), "feed_forward1")

        # ---- RelPositionMultiHeadAttention (update_cache, forward_qkv, rel_shift, forward_attention)
        a = ln(residual, "norm_self_att")
        B, Tq, _ = a.shape
        new_cache_ch = None
        if cache_ch is not None:
            kv_in = torch.cat([cache_ch, a], dim=1)
            keep = Tq - self.drop
            new_cache_ch = torch.cat([cache_ch[:, keep:, :], a[:, :keep, :]], dim=1)
        else:
            kv_in = a
        Tk = kv_in.size(1)
        q = self._lin(a, p + "self_attn.linear_q.weight").view(B, Tq, H, dk)
        k = self._lin(kv_in, p + "self_attn.linear_k.weight").view(B, Tk, H, dk).transpose(1, 2)
        v = self._lin(kv_in, p + "self_attn.linear_v.weight").view(B, Tk, H, dk).transpose(1, 2)
        pp = self._lin(pos_emb, p + "self_attn.linear_pos.weight").view(1, -1, H, dk).transpose(1, 2)
        q_u = (q + w[p + "self_attn.pos_bias_u"]).transpose(1, 2)
        q_v = (q + w[p + "self_attn.pos_bias_v"]).transpose(1, 2)
        if self.act_round == "bf16":
            q_u, q_v, k, v, pp = map(bf16_round, (q_u, q_v, k, v, pp))
        bd = torch.matmul(q_v, pp.transpose(-2, -1))
        # rel_shift
        b_, h_, ql, pl = bd.shape
        bd = F.pad(bd, pad=(1, 0)).view(b_, h_, -1, ql)[:, :, 1:].view(b_, h_, ql, pl)
        ac = torch.matmul(q_u, k.transpose(-2, -1))
        bd = bd[:, :, :, : ac.size(-1)]
        scores = (ac + bd) / math.sqrt(dk)
        m = att_mask.unsqueeze(1)
        scores = scores.masked_fill(m, -INF_VAL)
        attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
        if self.act_round == "bf16":
            attn = bf16_round(attn)
        ctx = torch.matmul(attn, v).transpose(1, 2).reshape(B, Tq, D)
        residual = residual + self._lin(ctx, p + "self_attn.linear_out.weight")

        # ---- ConformerConvolution (pw1, GLU, CausalConv1D w/ cache, BatchNorm eval, SiLU, pw2)
        c = ln(residual, "norm_conv")
        c = self._lin(c, p + "conv.pointwise_conv1.weight").transpose(1, 2)   # [B,2D,Tq]; Conv1d k=1 == Linear
        c = F.glu(c, dim=1)
        c = c.masked_fill(pad_mask.unsqueeze(1), 0.0)
        new_cache_tm = None
        if cache_tm is not None:
            new_x = torch.cat([cache_tm, F.pad(c, (0, 4))], dim=-1)
            nc = new_x[:, :, : -self.drop] if self.drop > 0 else new_x
            new_cache_tm = nc[:, :, -cache_tm.size(-1):]
        else:
            new_x = F.pad(c, (4, 4))
        dwv = F.conv1d(new_x, w[p + "conv.depthwise_conv.weight"], None, groups=D)
        dwv = F.batch_norm(dwv, w[p + "conv.batch_norm.running_mean"], w[p + "conv.batch_norm.running_var"],
                           w[p + "conv.batch_norm.weight"], w[p + "conv.batch_norm.bias"], False, 0.0, 1e-5)
        dwv = F.silu(dwv).transpose(1, 2)
        residual = residual + self._lin(dwv, p + "conv.pointwise_conv2.weight")

        residual = residual + 0.5 * ff(ln(residual, "norm_feed_forward2"), "feed_forward2")
        out = ln(residual, "norm_out")
        return out, new_cache_ch, new_cache_tm

    # ------------------------------------------------------------------ streaming step (contract layouts)
    @torch.no_grad()
    def stream_step(self, audio_signal, length, cache_last_channel, cache_last_time, cache_last_channel_len):
        """audio_signal [B,128,T] f32, length [B] i64, cache_last_channel [B,L,S,D], cache_last_time [B,L,D,K],
        cache_last_channel_len [B] i64 -> (encoder_output [B,D,<=3], encoded_lengths [B],
        cache_last_channel_out [B,L,S,D], cache_last_time_out [B,L,D,K], cache_last_channel_len_out [B])."""
        cc = cache_last_channel.transpose(0, 1)
        ct = cache_last_time.transpose(0, 1)
        x, length = self.pre_encode(audio_signal.transpose(1, 2), length)
        x = x[:, self.drop_pre:, :]
        length = (length - self.drop_pre).clamp(min=0)
        Tq = x.size(1)
        S = self.S
        keep = Tq - self.drop
        max_len = Tq + S
        padding_length = length + S
        offset = S - cache_last_channel_len
        pos_emb = self.rel_pos_emb(max_len)
        pad_mask, att_mask = self._create_masks(padding_length, max_len, offset)
        pad_mask = pad_mask[:, S:]
        att_mask = att_mask[:, S:]
        ncc, nct = [], []
        for i in range(self.L):
            x, c1, c2 = self._layer(i, x, att_mask, pos_emb, pad_mask, cc[i], ct[i])
            ncc.append(c1)
            nct.append(c2)
        enc = x.transpose(1, 2)
        ncc = torch.stack(ncc, 0)[:, :, -S:, :]
        nct = torch.stack(nct, 0)
        len_out = torch.clamp(cache_last_channel_len + keep, max=S)
        enc = enc[:, :, : self.valid_out]
        enc_len = torch.clamp(length, max=self.valid_out)
        return enc, enc_len, ncc.transpose(0, 1).contiguous(), nct.transpose(0, 1).contiguous(), len_out

    @torch.no_grad()
    def offline(self, audio_signal, length):
        """audio_signal [B,128,T] -> encoder_output [B,D,T_enc], encoded_lengths [B] (full attention)."""
        x, length = self.pre_encode(audio_signal.transpose(1, 2), length)
        T = x.size(1)
        pos_emb = self.rel_pos_emb(T)
        pad_mask, att_mask = self._create_masks(length, T, None)
        for i in range(self.L):
            x, _, _ = self._layer(i, x, att_mask, pos_emb, pad_mask, None, None)
        return x.transpose(1, 2), length

    def initial_cache(self, B: int):
        return (torch.zeros(B, self.L, self.S, self.D), torch.zeros(B, self.L, self.D, self.KT),
                torch.zeros(B, dtype=torch.int64))

    # ------------------------------------------------------------------ predictor / joint
    @torch.no_grad()
    def predictor_step(self, y: torch.Tensor, h: torch.Tensor, c: torch.Tensor):
        """y [B,1] i64, h,c [2,B,640] -> g [B,640,1], h_out, c_out  (RNNTDecoder.predict, add_sos=False)."""
        w = self.w
        x = F.embedding(y[:, 0], w["decoder.prediction.embed.weight"])
        hs, cs = [], []
        for l in range(self.cfg["pred_layers"]):
            pre = "decoder.prediction.dec_rnn.lstm."
            gates = (self._lin(x, pre + f"weight_ih_l{l}", pre + f"bias_ih_l{l}")
                     + self._lin(h[l], pre + f"weight_hh_l{l}", pre + f"bias_hh_l{l}"))
            i_, f_, g_, o_ = gates.chunk(4, dim=-1)          # PyTorch LSTM gate order i,f,g,o
            c_new = torch.sigmoid(f_) * c[l] + torch.sigmoid(i_) * torch.tanh(g_)
            h_new = torch.sigmoid(o_) * torch.tanh(c_new)
            hs.append(h_new)
            cs.append(c_new)
            x = h_new
        return x.unsqueeze(-1), torch.stack(hs, 0), torch.stack(cs, 0)

    @torch.no_grad()
    def joint_logits(self, enc_bdt: torch.Tensor, g_bhu: torch.Tensor) -> torch.Tensor:
        """encoder_output [B,1024,T], predictor_output [B,640,U] -> raw logits [B,T,U,8198]."""
        f = self._lin(enc_bdt.transpose(1, 2), "joint.enc.weight", "joint.enc.bias").unsqueeze(2)
        g = self._lin(g_bhu.transpose(1, 2), "joint.pred.weight", "joint.pred.bias").unsqueeze(1)
        return self._lin(F.relu(f + g), "joint.joint_net.2.weight", "joint.joint_net.2.bias")

    # ------------------------------------------------------------------ tokenizer predicates (cpp/src/tokenizer.cpp:25-84)
    def is_punct_only(self, tok: int) -> bool:
        if tok < 0 or tok >= len(self.vocab_lines):
            return False
        b = self.vocab_lines[tok].encode("utf-8")
        if b in (b"<blank>", b"<pad>", b"<unk>") or (len(b) > 0 and b[:1] == b"<" and b[-1:] == b">"):
            return False
        if b[:3] == b"\xe2\x96\x81":
            b = b[3:]
        if not b:
            return False
        any_non_space = False
        for ch in b:
            if (48 <= ch <= 57) or (65 <= ch <= 90) or (97 <= ch <= 122):
                return False
            if ch not in (9, 10, 11, 12, 13, 32):
                any_non_space = True
        return any_non_space


class DecodeState:
    """Per-stream predictor state carried across chunks (parakeet_trt.cpp:1595-1646)."""

    def __init__(self, model: ModelRef):
        self.h = torch.zeros(model.cfg["pred_layers"], 1, model.cfg["pred_hidden"])
        self.c = torch.zeros_like(self.h)
        self.g = None
        self.y_id = model.blank
        self.tokens: List[int] = []


def prime(model: ModelRef, st: DecodeState) -> None:
    """parakeet_reset_utterance priming (parakeet_trt.cpp:1886-1942): <|startoftranscript|> then <|en|>."""
    ids = {p: i for i, p in enumerate(model.vocab_lines)}
    for piece in ("<|startoftranscript|>", "<|en|>"):
        if piece in ids:
            st.g, st.h, st.c = model.predictor_step(torch.tensor([[ids[piece]]]), st.h, st.c)
            st.y_id = ids[piece]
    if st.g is None:  # tdt_trace.py:233-235
        st.g, st.h, st.c = model.predictor_step(torch.tensor([[model.blank]]), st.h, st.c)


def tdt_greedy_chunk(model: ModelRef, st: DecodeState, enc_out: torch.Tensor, t_enc: int,
                     max_symbols: int = 8, punct_suppression: bool = True,
                     margins: Optional[list] = None, blank_penalty: float = 0.0) -> List[Tuple[int, int, int, int]]:
    """Greedy TDT over one chunk's encoder frames; mutates st.  Returns the per-step trace
    [(time_idx, best_tok, duration, advance)].  Follows parakeet_trt.cpp:2914-3676 / tdt_trace.py:277-353:
    first-max-wins argmax (strict '>'), blank+dur0 -> advance 1, non-blank -> predictor step,
    advance 0 -> stay, forced +1 after max_symbols, leftover advance dropped at chunk end,
    leading punctuation-only suppression while nothing has been emitted (:3256-3262); NaN logits read as -100 (:2971, both heads);
    PARAKEET_BLANK_PENALTY is subtracted from the blank logit before the token argmax (:3175-3178)."""
    dur_values = [0, 1, 2, 3, 4]
    V = model.vocab
    trace = []
    t = 0
    while t < t_enc:
        advanced = False
        for _u in range(max_symbols):
            logits = model.joint_logits(enc_out[:, :, t:t + 1], st.g)[0, 0, 0]
            logits = torch.where(torch.isnan(logits), torch.full_like(logits, -100.0), logits)
            if blank_penalty != 0.0:
                logits = logits.clone()
                logits[model.blank] -= blank_penalty
            tok = int(torch.argmax(logits[:V]))       # torch.argmax returns the first maximal index
            if punct_suppression and not st.tokens and model.is_punct_only(tok):
                tok = model.blank
            d = dur_values[int(torch.argmax(logits[V:V + model.n_dur]))]
            if margins is not None:     # top-2 gaps of both heads: how far this decision is from flipping
                t2 = torch.topk(logits[:V], 2).values
                d2 = torch.topk(logits[V:V + model.n_dur], 2).values
                margins.append((float(t2[0] - t2[1]), float(d2[0] - d2[1])))
            adv = 1 if (tok == model.blank and d == 0) else d
            trace.append((t, tok, d, adv))
            if tok != model.blank:
                st.tokens.append(tok)
                st.g, st.h, st.c = model.predictor_step(torch.tensor([[tok]]), st.h, st.c)
                st.y_id = tok
            if adv == 0:
                continue
            t += adv
            advanced = True
            break
        if not advanced:
            t += 1
    return trace


def streaming_schedule(num_chunks: int, chunk_size=(41, 48), shift_size=(17, 24), pre_encode=(0, 9)):
    """streaming_encoder_reference.py:522-550.  Returns [(slice_start, slice_end)] in feature frames."""
    out, start = [], 0
    for idx in range(num_chunks):
        r = 0 if idx == 0 else 1
        out.append((max(0, start - pre_encode[r]), start + chunk_size[r]))
        start += shift_size[r]
    return out


def decode_text(vocab_lines: List[str], ids: List[int]) -> str:
    """cpp/src/tokenizer.cpp:32-57."""
    out = ""
    for i in ids:
        if i < 0 or i >= len(vocab_lines):
            continue
        tok = vocab_lines[i]
        if tok in ("<blank>", "<pad>", "<unk>") or (tok and tok[0] == "<" and tok[-1] == ">"):
            continue
        if tok.startswith("▁"):
            if out and not out.endswith(" "):
                out += " "
            out += tok[1:]
            continue
        out += tok
    return out.lstrip(" ")
