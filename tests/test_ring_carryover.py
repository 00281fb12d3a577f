"""Cache carry-over on the CPU: the engine keeps projected K / V rows in a 288-slot RING addressed by
a per-stream head instead of shifting the reference's FIFO tensor cache_last_channel [256 rows] every chunk
(parakeet_trt.cpp:2547-2755; NeMo update_cache: new_cache = cat(cache[keep:], rows[:keep]), keep = Tq - cache_drop_size).
This test restates the ring bookkeeping of csrc/engine.cu (run_batch: head += keep, len = min(len + keep, 256)), the
physical-order key walk and validity rule of csrc/attn_mma.cu and the position index (256 + i - j) in numpy, and checks it chunk
by chunk -- through several wrap-arounds and with mixed chunk sizes -- against the FIFO formulation with NeMo's masks and
rel_shift (oracle/model_ref.py:_layer / stream_step).  It pins the ALGORITHM; the CUDA code is checked on the GPU."""
import math

import numpy as np
import torch
import torch.nn.functional as F

S, CAP, DROP, DK, MAXTQ = 256, 288, 3, 16, 32
POS_NEG = MAXTQ - 1                       # table row of relative position r is r + POS_NEG (common.cuh kPosNeg)


def fifo_attention(cache, cache_len, rows, table_rel):
    """NeMo formulation: keys = cat(cache[256], rows[Tq]); mask = positions < 256 - cache_len; rel_shift over 2*(256+Tq)-1 rows."""
    Tq = rows.shape[0]
    kv = np.concatenate([cache, rows], 0)
    L = S + Tq
    pp = np.stack([table_rel(rel) for rel in range(L - 1, -L, -1)])            # positions L-1 ... -(L-1)
    q = torch.from_numpy(rows)[None, None]
    # full-length query axis as in NeMo (queries padded to L, the last Tq rows are the real ones)
    qfull = torch.zeros(1, 1, L, DK, dtype=torch.float64)
    qfull[0, 0, S:] = q[0, 0]
    bd = qfull @ torch.from_numpy(pp).T[None, None]
    b_, h_, ql, pl = bd.shape
    bd = F.pad(bd, pad=(1, 0)).view(b_, h_, -1, ql)[:, :, 1:].view(b_, h_, ql, pl)[..., :L]
    ac = qfull @ torch.from_numpy(kv).T[None, None]
    sc = ((ac + bd) / math.sqrt(DK))[0, 0, S:]                                  # [Tq, L]
    mask = torch.arange(L) < (S - cache_len)
    sc = sc.masked_fill(mask[None, :], -1e4)
    att = torch.softmax(sc, dim=-1).masked_fill(mask[None, :], 0.0)
    return (att @ torch.from_numpy(kv)).numpy()


def ring_attention(ring, head, cache_len, Tq, table_rel):
    """Engine formulation: walk the 288 physical slots; slot p holds logical position j = (p - head) mod 288."""
    out = np.zeros((Tq, DK))
    q = np.stack([ring[(head + S + i) % CAP] for i in range(Tq)])
    for i in range(Tq):
        sc = np.full(CAP, -np.inf)
        for p in range(CAP):
            j = (p - head) % CAP
            if S - cache_len <= j < S + Tq:
                sc[p] = (q[i] @ ring[p] + q[i] @ table_rel(S + i - j)) / math.sqrt(DK)
        w = np.exp(sc - sc.max())
        w /= w.sum()
        out[i] = w @ ring
    return out


def test_ring_equals_fifo_over_wraps():
    rng = np.random.default_rng(7)
    table = rng.standard_normal((S + 2 * MAXTQ, DK))                              # rows for rel in [-POS_NEG, 256 + 32]
    table_rel = lambda rel: table[rel + POS_NEG] if -POS_NEG <= rel < S + MAXTQ + 1 else np.zeros(DK)
    fifo = np.zeros((S, DK)); fifo_len = 0
    ring = np.zeros((CAP, DK)); head = 0; ring_len = 0
    sizes = [4] + [6] * 60 + [15, 6, 30, 6, 6, 4] + [6] * 60                      # > 3 wrap-arounds of the 288-slot ring
    for n, Tq in enumerate(sizes):
        rows = rng.standard_normal((Tq, DK))
        for i in range(Tq):                                                        # the QKV epilogue's scatter
            ring[(head + S + i) % CAP] = rows[i]
        got = ring_attention(ring, head, ring_len, Tq, table_rel)
        want = fifo_attention(fifo, fifo_len, rows, table_rel)
        assert np.abs(got - want).max() < 1e-9, (n, Tq)
        keep = Tq - DROP
        fifo = np.concatenate([fifo[keep:], rows[:keep]], 0)                       # NeMo update_cache
        fifo_len = min(fifo_len + keep, S)
        head = (head + keep) % CAP                                                  # Engine::run_batch
        ring_len = min(ring_len + keep, S)
        assert ring_len == fifo_len
    assert fifo_len == S and head != 0
