"""The whole-utterance attention kernel (csrc/offline_long.cu) never forms NeMo's [T, 2T-1] position-score matrix: per 64 x 64
score tile it multiplies the queries with the 127-row WINDOW of the projected table the tile can touch and applies rel_shift as
the index skew  S[r][c] += G[r][r - c + 63]  (G over window rows  i0 - j0 - 63 ... i0 - j0 + 63,  table row = rel + Tm - 1).
This CPU test restates exactly that tiling in numpy (including the per-warp 80-column sub-window and the online softmax) and
checks it against the oracle's pad / view / slice rel_shift formulation (oracle/model_ref.py:_layer), for ragged lengths and a
table built for a longer utterance.  It pins the ALGORITHM of the kernel; the CUDA code itself is checked on the GPU
(tests/test_gpu_offline_long.py)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

BM = BN = 64
DK = 128


def oracle_attention(q_u, q_v, k, v, pp):
    """q_u, q_v, k, v: [T, dk]; pp: [2T-1, dk] rows for relative positions T-1 ... -(T-1)  (NeMo order)."""
    T = q_u.shape[0]
    bd = torch.from_numpy(q_v @ pp.T)[None, None]                       # [1,1,T,2T-1]
    b_, h_, ql, pl = bd.shape
    bd = F.pad(bd, pad=(1, 0)).view(b_, h_, -1, ql)[:, :, 1:].view(b_, h_, ql, pl)[..., :T]
    ac = torch.from_numpy(q_u @ k.T)[None, None]
    attn = torch.softmax((ac + bd) / math.sqrt(DK), dim=-1)
    return (attn @ torch.from_numpy(v))[0, 0].numpy()


def tiled_attention(q_u, q_v, k, v, table, Tm):
    """table: [2Tm-1, dk], row = rel + Tm - 1 (ascending relative position, as lf_posemb_kernel builds it)."""
    T = q_u.shape[0]
    out = np.zeros((T, DK), np.float64)
    for i0 in range(0, T, BM):
        rows = min(BM, T - i0)
        m = np.full(BM, -np.inf)
        l = np.zeros(BM)
        o = np.zeros((BM, DK))
        qu = np.zeros((BM, DK)); qu[:rows] = q_u[i0:i0 + rows]
        qv = np.zeros((BM, DK)); qv[:rows] = q_v[i0:i0 + rows]
        for j0 in range(0, T, BN):
            kt = np.zeros((BN, DK)); vt = np.zeros((BN, DK))
            n = min(BN, T - j0)
            kt[:n] = k[j0:j0 + n]; vt[:n] = v[j0:j0 + n]
            win = np.zeros((128, DK))                                        # window rows outside the table read as zero (TMA)
            for r in range(128):
                prow = i0 - j0 - (BN - 1) + r + (Tm - 1)
                if 0 <= prow < 2 * Tm - 1:
                    win[r] = table[prow]
            s = qu @ kt.T                                                    # [64, 64]
            for w in range(4):                                               # warp w: rows 16w..16w+15, its 80 window columns
                g = qv[16 * w:16 * w + 16] @ win[16 * w:16 * w + 80].T       # [16, 80]
                for rl in range(16):
                    for c in range(BN):
                        s[16 * w + rl, c] += g[rl, rl - c + (BN - 1)]
            s = s / math.sqrt(DK)
            s[:, n:] = -np.inf                                               # keys past the utterance
            m_new = np.maximum(m, s.max(axis=1))
            alpha = np.exp(m - m_new)
            p = np.exp(s - m_new[:, None])
            l = l * alpha + p.sum(axis=1)
            o = o * alpha[:, None] + p @ vt
            m = m_new
        out[i0:i0 + rows] = (o / l[:, None])[:rows]
    return out


@pytest.mark.parametrize("T,Tm", [(1, 1), (7, 7), (64, 64), (65, 65), (130, 130), (100, 333)], ids=str)
def test_windowed_skew_equals_rel_shift(T, Tm):
    rng = np.random.default_rng(T * 1000 + Tm)
    q = rng.standard_normal((T, DK))
    u, vb = 0.3 * rng.standard_normal(DK), 0.3 * rng.standard_normal(DK)
    k, v = rng.standard_normal((T, DK)), rng.standard_normal((T, DK))
    table = rng.standard_normal((2 * Tm - 1, DK))                            # row = rel + Tm - 1, rel = -(Tm-1) ... Tm-1
    # NeMo's order for an utterance of length T: relative positions T-1 ... -(T-1)
    pp = table[[rel + Tm - 1 for rel in range(T - 1, -T, -1)]]
    want = oracle_attention(q + u, q + vb, k, v, pp)
    got = tiled_attention(q + u, q + vb, k, v, table, Tm)
    assert np.abs(got - want).max() < 1e-9
