#!/usr/bin/env python3
"""Bring-up helper: runs one named section on the GPU box and prints error summaries (used under `timeout`)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402  (sets sys.path)
import torch  # noqa: E402

import binding  # noqa: E402
from model_ref import DecodeState, ModelRef, prime, streaming_schedule, tdt_greedy_chunk  # noqa: E402
from synth_audio import synth_clip  # noqa: E402
from weights_io import bf16_bits_to_f32, f32_to_bf16_bits  # noqa: E402


def feats(fr, seconds, seed):
    f = conftest.normalized_features(fr, seconds, seed)
    f[0] = 0
    return f


def main():
    sec = sys.argv[1]
    layers = int(os.environ.get("LAYERS", "2"))
    prec = int(os.environ.get("PREC", "1"))
    backend = int(os.environ.get("BACKEND", "1"))
    model = conftest.model_dir(layers)
    fr = conftest.FeaturesRef(conftest.build_oracle())
    pre = os.environ.get("PRE", "")
    if pre:
        for pp in pre.split(","):
            e0 = binding.Engine(model, max_streams=4 if pp[0] == "f" else 1, precision=int(pp[-1]))
            if pp[0] == "g":
                rng = np.random.default_rng(1)
                for M, N, K in [(1, 256, 256), (300, 1024, 1024), (64, 8198, 640), (1000, 3072, 1024)]:
                    A = rng.standard_normal((M, K)).astype(np.float32)
                    Wb = f32_to_bf16_bits(rng.standard_normal((N, K)).astype(np.float32)).reshape(N, K)
                    e0.gemm_test(0, A, Wb); e0.gemm_test(1, A, Wb)
            else:
                e0.logmel(synth_clip(10.0, 1)); e0.logmel(synth_clip(60.0, 2), True)
            e0.close()
            print("  pre-engine", pp, "done", flush=True)
    t0 = time.time()
    eng = binding.Engine(model, max_streams=int(os.environ.get("STREAMS", "4")), precision=prec, gemm_backend=backend, max_rows=int(os.environ.get("ROWS", "64")))
    print(f"[{sec}] engine up in {time.time()-t0:.1f}s layers={layers} prec={prec} backend={backend}", flush=True)
    if sec == "frontend":
        for s in (0.5, 10.0):
            pcm = synth_clip(s, 1234)
            d = np.abs(eng.logmel(pcm) - fr.logmel(pcm))
            print(f"  logmel {s}s max|err|={d.max():.3e} mean={d.mean():.3e}")
        z = eng.logmel(np.zeros(16000, np.float32))
        print("  zeros exact:", bool(np.all(z == np.float32(np.log(np.float32(1e-5))))), z[0, 0])
    elif sec == "gemm":
        for M, N, K in [(96, 256, 256), (352, 256, 256), (256, 2048, 1024), (6, 1024, 1024), (300, 1024, 4096), (64, 8198, 640)]:
            rng = np.random.default_rng(1)
            A = rng.standard_normal((M, K)).astype(np.float32)
            Wb = f32_to_bf16_bits(rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K)).reshape(N, K)
            W = bf16_bits_to_f32(Wb).reshape(N, K)
            Ae = A if prec == 1 else bf16_bits_to_f32(f32_to_bf16_bits(A)).reshape(M, K)
            ref = Ae.astype(np.float64) @ W.astype(np.float64).T
            for be in (0, 1):
                C = eng.gemm_test(be, A, Wb)
                print(f"  gemm M{M} N{N} K{K} backend={be} max|err|={np.max(np.abs(C-ref)):.3e} ref_rms={np.sqrt((ref**2).mean()):.3f}", flush=True)
    elif sec == "encoder":
        m = ModelRef(model)
        f = feats(fr, 3.0, 1234)
        cc, ct, cl = m.initial_cache(1)
        gcc, gct, gcl = cc.numpy().copy(), ct.numpy().copy(), cl.numpy().copy()
        for k, (b, e) in enumerate(streaming_schedule(int(os.environ.get("CHUNKS", "4")))):
            x = f[None, :, b:e]
            enc, el, cc, ct, cl = m.stream_step(torch.from_numpy(x), torch.tensor([e - b]), cc, ct, cl)
            genc, gel, gcc, gct, gcl = eng.encoder_streaming_step(x, np.array([e - b]), gcc, gct, gcl)
            d = np.abs(genc - enc.numpy())
            print(f"  chunk {k}: enc max|err|={d.max():.3e} p95={np.percentile(d,95):.3e} |enc|max={enc.abs().max():.2f} "
                  f"cache_ch err={np.abs(gcc-cc.numpy()).max():.3e} cache_tm err={np.abs(gct-ct.numpy()).max():.3e} len {gcl.tolist()} {cl.tolist()}", flush=True)
    elif sec == "saturated":
        m = ModelRef(model)
        x = feats(fr, 1.0, 9)[None, :, :57]
        for trial, (ln, scale) in enumerate([(256, 1.0), (256, 1.0), (256, 0.0), (100, 1.0), (16, 1.0), (256, 0.2)]):
            rng = np.random.default_rng(5)
            cc = (scale * rng.standard_normal((1, m.L, 256, 1024))).astype(np.float32)
            ct = (scale * rng.standard_normal((1, m.L, 1024, 4))).astype(np.float32)
            cc[:, :, :256 - ln] = 0
            enc, el, cco, cto, clo = m.stream_step(torch.from_numpy(x), torch.tensor([57]), torch.from_numpy(cc), torch.from_numpy(ct), torch.tensor([ln]))
            genc, gel, gcc, gct, gcl = eng.encoder_streaming_step(x, np.array([57]), cc, ct, np.array([ln]))
            d = np.abs(genc - enc.numpy())
            print(f"  trial {trial} len={ln} scale={scale}: enc max|err|={d.max():.3e} p95={np.percentile(d,95):.3e} cache_ch err={np.abs(gcc-cco.numpy()).max():.3e} "
                  f"cache_tm err={np.abs(gct-cto.numpy()).max():.3e}", flush=True)
    elif sec == "lensweep":
        m = ModelRef(model)
        for T in (57, 41):
            x = feats(fr, 1.0, 9)[None, :, :T]
            for ln in (100, 120, 127, 128, 129, 134, 135, 140, 156, 200, 255, 256):
                for scale in (0.0, 1.0):
                    rng = np.random.default_rng(5)
                    cc = (scale * rng.standard_normal((1, m.L, 256, 1024))).astype(np.float32)
                    ct = (scale * rng.standard_normal((1, m.L, 1024, 4))).astype(np.float32)
                    cc[:, :, :256 - ln] = 0
                    enc, el, cco, cto, clo = m.stream_step(torch.from_numpy(x), torch.tensor([T]), torch.from_numpy(cc), torch.from_numpy(ct), torch.tensor([ln]))
                    genc, gel, gcc, gct, gcl = eng.encoder_streaming_step(x, np.array([T]), cc, ct, np.array([ln]))
                    d = np.abs(genc - enc.numpy())
                    tm = np.abs(gct - cto.numpy())
                    print(f"  T={T} len={ln} scale={scale}: enc max|err|={d.max():.3e} cache_tm err L0={tm[0,0].max():.3e} L1={tm[0,1].max():.3e}", flush=True)
    elif sec == "decode":
        m = ModelRef(model)
        torch.manual_seed(0)
        y, h, c = torch.tensor([[17], [8192], [4000]]), 0.5 * torch.randn(2, 3, 640), torch.randn(2, 3, 640)
        g, ho, co = m.predictor_step(y, h, c)
        gg, gh, gc = eng.predictor_step(y.numpy(), h.numpy(), c.numpy())
        print(f"  predictor g err={np.abs(gg-g.numpy()).max():.3e} h err={np.abs(gh-ho.numpy()).max():.3e} c err={np.abs(gc-co.numpy()).max():.3e}")
        enc = torch.randn(3, 1024, 2)
        lg = m.joint_logits(enc, g).numpy()
        glg = eng.joint_step(enc.numpy(), g.numpy())
        print(f"  joint logits err={np.abs(glg-lg).max():.3e} argmax eq={np.array_equal(glg[...,:8193].argmax(-1), lg[...,:8193].argmax(-1))}")
    elif sec == "stream":
        m = ModelRef(model)
        n_streams, n_chunks = int(os.environ.get("NS", "3")), int(os.environ.get("CHUNKS", "8"))
        fl = [feats(fr, 1.0 + 0.25 * n_chunks, 1000 + i) for i in range(n_streams)]
        sids = [eng.open() for _ in fl]
        ora = []
        for _ in fl:
            st = DecodeState(m); prime(m, st); ora.append([st, *m.initial_cache(1)])
        same = tot = 0
        for k, (b, e) in enumerate(streaming_schedule(n_chunks)):
            for i, s in enumerate(sids):
                eng.push_features(s, fl[i][:, b:e])
            t1 = time.time(); eng.step(); dt = time.time() - t1
            for i, s in enumerate(sids):
                st, cc, ct, cl = ora[i]
                enc, el, cc, ct, cl = m.stream_step(torch.from_numpy(fl[i][None, :, b:e]), torch.tensor([e - b]), cc, ct, cl)
                ora[i][1:] = [cc, ct, cl]
                want = [(t, tok, d) for t, tok, d, _ in tdt_greedy_chunk(m, st, enc, int(el))]
                got = eng.last_steps(s)
                tot += 1; same += int(got == want)
                if got != want:
                    print(f"  MISMATCH chunk {k} stream {i}: got {got} want {want}")
            print(f"  chunk {k}: step {dt*1e3:.2f} ms", flush=True)
        print(f"  identical chunks {same}/{tot}; tokens stream0: {eng.tokens(sids[0])}")
        print("  text:", eng.text(sids[0])[:100])
    print(f"[{sec}] done, launches={eng.kernel_launches()}", flush=True)


if __name__ == "__main__":
    main()
