"""Helper process of tests/test_gpu_ab_switches.py: runs the frontend and a staggered batch of streaming streams with whatever kernel
switches the environment selects (the switches are read once per process) and writes everything observable to an .npz file.
usage: _ab_driver.py <package dir> <model dir> <out.npz> <n_streams> <n_chunks>"""
import sys

import numpy as np

sys.path.insert(0, sys.argv[1])
import binding  # noqa: E402
from synth_audio import synth_clip  # noqa: E402


def main():
    model, out, n_streams, n_chunks = sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    eng = binding.Engine(model, max_streams=n_streams, precision=0)
    res = {}
    for i, (sec, seed) in enumerate([(10.0, 1234), (0.025, 5), (3.37, 3)]):
        res[f"logmel{i}"] = eng.logmel(synth_clip(sec, seed))
    res["logmel_norm"] = eng.logmel(synth_clip(10.0, 1234), per_feature_norm=True)
    # streams start 5 chunks apart: every step mixes cache lengths from 0 to saturated; n_chunks > 96 wraps the 288-slot rings
    clips = [synth_clip(0.41 + 0.24 * n_chunks + 0.3, 900 + i) for i in range(4)]
    sids = [eng.open() for _ in range(n_streams)]
    first = 400 + 40 * 160      # 41 frames: the first chunk; every later chunk consumes 24 new frames
    hop = 24 * 160
    traces = []
    lens = []
    for c in range(n_chunks + 5 * n_streams):
        for i, s in enumerate(sids):
            k = c - 5 * i
            if k < 0 or k >= n_chunks:
                continue
            pcm = clips[i % 4]
            lo = 0 if k == 0 else first + (k - 1) * hop
            hi = first + k * hop
            eng.push_audio(s, pcm[lo:hi])
        while eng.step():
            traces.append([x for s in sids for t in eng.last_steps(s) for x in t])
        lens.append([eng.cache_len(s) for s in sids])
    res["traces"] = np.array([x for t in traces for x in t + [-7]], np.int64)
    res["lens"] = np.array(lens, np.int64)
    res["tokens"] = np.array([x for s in sids for x in eng.tokens(s) + [-1]], np.int64)
    for i in (0, n_streams - 1):
        st = eng.export_state(sids[i])
        res[f"state{i}_ch"] = np.asarray(st[0])
        res[f"state{i}_tm"] = np.asarray(st[1])
    eng.close()
    np.savez(out, **res)
    print("AB-OK")


main()
