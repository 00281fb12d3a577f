"""The six legacy entry points (include/parakeet_trt.h) beyond the basic push / poll / reset test in test_gpu_model.py:
  * the push prologue: a push longer than 256 frames (or than PARAKEET_MAX_FRAMES_PER_PUSH) is re-sliced into maximal pushes
    (/root/reference/cpp/src/parakeet_trt.cpp:1982-2011) -- decode traces identical to the same slices pushed by hand;
  * the PARAKEET_DEBUG_TDT_STEPS trace of THIS build, regenerated live, against the reference's own PyTorch-loop trace
    (tests/golden/tdt_trace_ref.json) -- and through the reference's own compare_tdt_trace.py where that tree exists;
  * several sessions share one engine: memory does not grow per session, concurrent pushes coalesce, traces equal single-session runs;
  * NaN guard toggles (parakeet_trt.cpp:913-1013)."""
import json
import os
import re
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

import binding
from conftest import ROOT, normalized_features
from model_ref import DecodeState, prime, streaming_schedule, tdt_greedy_chunk

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(ROOT, "tests", "_legacy_driver.py")
RX = re.compile(r"tdt_step time_idx=(\d+) u=(\d+) best_tok=(\d+) best_dur_idx=(\d+) duration=(\d+) advance=(\d+) blank=(\d) blank_dur0_clamped=(\d)")


def _drive(model, fpath, pushes, fp16=False, **env):
    e = dict(os.environ, PARAKEET_DEBUG_TDT_STEPS="1000000", **{k: str(v) for k, v in env.items()})
    run = subprocess.run([sys.executable, DRIVER, model, str(fpath), "1" if fp16 else "0"] + [f"{a}:{b}" for a, b in pushes],
                         capture_output=True, text=True, env=e, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    rcs = [int(ln[3:]) for ln in run.stdout.splitlines() if ln.startswith("rc=")]
    events = [ln for ln in run.stdout.splitlines() if ln.startswith("event ")]
    steps = [tuple(int(x) for x in m.groups()) for m in map(RX.search, run.stderr.splitlines()) if m]
    return rcs, steps, events, run.stderr


def _save_feats(tmp_path, features_ref, seconds, seed):
    f = normalized_features(features_ref, seconds, seed)
    f[0] = 0.0
    np.save(tmp_path / "f.npy", f)
    return f, tmp_path / "f.npy"


def test_push_prologue_auto_chunking(tmp_path, model_small, features_ref):
    f, fp = _save_feats(tmp_path, features_ref, 7.0, 31)
    assert f.shape[1] >= 600
    # 600 frames in ONE push == 256 + 256 + 88 by hand (the reference CLI slices at 256 itself, the ABI must do it too)
    rc1, st1, ev1, _ = _drive(model_small, fp, [(0, 600)], PARAKEET_EMIT_FINAL_EACH_CHUNK=1)
    rc2, st2, ev2, _ = _drive(model_small, fp, [(0, 256), (256, 512), (512, 600)], PARAKEET_EMIT_FINAL_EACH_CHUNK=1)
    assert rc1 == [0] and rc2 == [0, 0, 0]
    assert len(st1) >= 3 and st1 == st2
    assert [e for e in ev1 if e.startswith("event final")] == [e for e in ev2 if e.startswith("event final")] and len(ev1) >= 3
    # 257 frames: 256 + a 1-frame tail the streaming encoder cannot take -> the first slice is processed, then rc -2 + ERROR event
    rc3, st3, ev3, _ = _drive(model_small, fp, [(0, 257)])
    rc4, st4, _, _ = _drive(model_small, fp, [(0, 256), (256, 257)])
    assert rc3 == [-2] and rc4 == [0, -2] and st3 == st4 and len(st3) >= 1
    assert any(e.startswith("event error") and "frames" in e for e in ev3)
    # PARAKEET_MAX_FRAMES_PER_PUSH=100 (:1984): 600 frames -> six 100-frame slices
    rc5, st5, _, _ = _drive(model_small, fp, [(0, 600)], PARAKEET_MAX_FRAMES_PER_PUSH=100)
    rc6, st6, _, _ = _drive(model_small, fp, [(i, i + 100) for i in range(0, 600, 100)])
    assert rc5 == [0] and rc6 == [0] * 6 and st5 == st6 and st5 != st1
    # a push at the limit is not re-sliced
    rc7, st7, _, _ = _drive(model_small, fp, [(0, 256)])
    assert rc7 == [0] and st7 == st2[:len(st7)]


def test_live_trace_matches_the_references_python_loop(tmp_path, model_small, features_ref):
    """Regenerates the PARAKEET_DEBUG_TDT_STEPS trace on THIS box (offline encoder mode, fp32-grade arithmetic, the clip of
    tests/golden/tdt_trace_ref.json = the trace of the reference's tools/verify_nemo/tdt_trace.py executed unmodified) and compares step
    by step; where /root/reference exists the reference's own triage tool does the comparison as well."""
    doc = json.load(open(os.path.join(ROOT, "tests", "golden", "tdt_trace_ref.json")))
    f, fp = _save_feats(tmp_path, features_ref, doc["clip"]["seconds"], doc["clip"]["seed"])
    C = doc["chunk_frames"]
    pushes = [(lo, min(lo + C, f.shape[1])) for lo in range(0, f.shape[1], C)]
    rcs, steps, _, stderr = _drive(model_small, fp, pushes, PARAKEET_B200_ENCODER="offline", PARAKEET_DISABLE_PUNCT_SUPPRESSION=1)
    assert all(rc == 0 for rc in rcs)
    want = [(s_["time_idx"], s_["u"], s_["best_tok"], s_["best_dur_idx"], s_["duration"], s_["advance"], int(s_["best_tok"] == 8192),
             int(s_["best_tok"] == 8192 and s_["duration"] == 0)) for s_ in doc["steps"]]
    amb = next((i for i, s_ in enumerate(doc["steps"]) if min(s_["tok_gap"], s_["dur_gap"]) < 1e-3), len(want))
    assert amb >= 60 and steps[:amb] == want[:amb]
    if amb == len(want):
        assert steps == want
    tool = "/root/reference/tools/verify_nemo/compare_tdt_trace.py"
    if os.path.exists(tool) and amb == len(want):
        pt, log = tmp_path / "pt.jsonl", tmp_path / "cpp.log"
        log.write_text(stderr)
        with open(pt, "w") as fh:
            fh.write(json.dumps(doc["meta"]) + "\n")
            for i, s_ in enumerate(doc["steps"]):
                blank = s_["best_tok"] == doc["meta"]["blank_id"]
                fh.write(json.dumps({"type": "step", "step_idx": i, **{k: s_[k] for k in ("chunk_idx", "time_idx", "u", "best_tok", "best_dur_idx", "duration", "advance")},
                                     "is_blank": blank, "blank_dur0_clamped": bool(blank and s_["duration"] == 0)}) + "\n")
        run = subprocess.run([sys.executable, tool, "--pt-trace", str(pt), "--cpp-stderr", str(log), "--check-index", "--fields",
                              "best_tok,best_dur_idx,duration,advance,is_blank,blank_dur0_clamped"], capture_output=True, text=True)
        assert run.returncode == 0 and f"matched {len(doc['steps'])} steps" in run.stdout, run.stdout + run.stderr


def test_sessions_share_one_engine(model_full, features_ref):
    """8 legacy sessions of one process: ONE set of weights (the first session pays for engine + slots, each further one < 50 MB),
    pushes from 8 threads are served by shared batched passes, every session's text equals a single-session run of its clip."""
    n = 8
    feats = []
    for i in range(n):
        f = normalized_features(features_ref, 3.2, 500 + i)
        f[0] = 0.0
        feats.append(f)
    sched = streaming_schedule(10)

    def run_one(sess, f, out):
        for b, e in sched:
            sess.push_features(f[:, b:e], e - b)
        # final transcript through the event the session emits for its last chunk (PARAKEET_EMIT_FINAL_EACH_CHUNK) is per chunk;
        # read the accumulated text from a last PARTIAL instead: force one by waiting out the 100 ms rate limit
        out.append(True)

    os.environ["PARAKEET_EMIT_FINAL_EACH_CHUNK"] = "1"
    try:
        # single-session baseline: one session alive at a time
        base = []
        for i in range(n):
            s = binding.ParakeetSessionSafe(model_full, 0, use_fp16=False)
            finals = []
            for b, e in sched:
                s.push_features(feats[i][:, b:e], e - b)
                while (ev := s.poll_event()) is not None:
                    if ev.kind == "final":
                        finals.append(ev.text)
            base.append(finals)
            s.close()
        torch.cuda.synchronize()
        free0, _ = torch.cuda.mem_get_info()
        first = binding.ParakeetSessionSafe(model_full, 0, use_fp16=False)
        free1, _ = torch.cuda.mem_get_info()
        rest = [binding.ParakeetSessionSafe(model_full, 0, use_fp16=False) for _ in range(n - 1)]
        free2, _ = torch.cuda.mem_get_info()
        per_extra = (free1 - free2) / (n - 1)
        print(f"\n[shared engine] first session {(free0 - free1) >> 20} MiB, each further session {per_extra / 2**20:.1f} MiB")
        assert per_extra < 50 * 2**20
        sessions = [first] + rest
        got = [[] for _ in range(n)]
        errs = []

        def worker(i):
            try:
                for b, e in sched:
                    sessions[i].push_features(feats[i][:, b:e], e - b)
                    while (ev := sessions[i].poll_event()) is not None:
                        if ev.kind == "final":
                            got[i].append(ev.text)
            except Exception as ex:      # noqa: BLE001
                errs.append((i, repr(ex)))
        th = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs
        assert got == base and any(any(x for x in g) for g in got)
        # a ninth session does not fit the 8 slots of the shared engine: a second engine is created for it, and it still works
        ninth = binding.ParakeetSessionSafe(model_full, 0, use_fp16=False)
        ninth.push_features(feats[0][:, :41], 41)
        ninth.close()
        for s in sessions:
            s.close()
        torch.cuda.synchronize()
        free3, _ = torch.cuda.mem_get_info()
        assert free0 - free3 < 64 << 20, "closing the last session must release the shared engine"
    finally:
        del os.environ["PARAKEET_EMIT_FINAL_EACH_CHUNK"]


def test_nan_guard_toggles(tmp_path, model_small, features_ref):
    """A NaN planted in a LayerNorm bias poisons encoder_output and the caches: the guard reports it in the reference's line format;
    PARAKEET_NAN_GUARD_HALT aborts the process; a clean model stays silent; after the first 10 guarded tensors only 1 in 100 is checked
    unless PARAKEET_NAN_GUARD_ALWAYS is set."""
    import shutil
    import struct

    from test_gpu_decode_extras import _tensor_offset
    f, fp = _save_feats(tmp_path, features_ref, 4.0, 8)
    d = tmp_path / "nanmodel"
    d.mkdir()
    shutil.copyfile(os.path.join(model_small, "weights.bin"), d / "weights.bin")
    shutil.copyfile(os.path.join(model_small, "vocab.txt"), d / "vocab.txt")
    off, nb, dt = _tensor_offset(str(d / "weights.bin"), "encoder.layers.1.norm_out.bias")
    assert dt == 0
    with open(d / "weights.bin", "r+b") as fh:
        fh.seek(off + 4 * 5)
        fh.write(struct.pack("<f", float("nan")))
    pushes = [(b, e) for b, e in streaming_schedule(12)]
    _, _, _, clean = _drive(model_small, fp, pushes)
    assert "NAN_GUARD" not in clean
    rcs, _, _, err = _drive(str(d), fp, pushes)
    alerts = [ln for ln in err.splitlines() if "NAN_GUARD ALERT" in ln]
    assert rcs == [0] * 12 and alerts and "stage=enc_output" in alerts[0] and "nan_count=" in alerts[0] and "first_nan_idx=" in alerts[0]
    n_default = len(alerts)
    _, _, _, err = _drive(str(d), fp, pushes, PARAKEET_NAN_GUARD_ALWAYS=1)
    n_always = sum("NAN_GUARD ALERT" in ln for ln in err.splitlines())
    assert n_always >= 12 and n_always > n_default      # sampling: 10 tensors, then 1 in 100
    e = dict(os.environ, PARAKEET_NAN_GUARD_HALT="1")
    run = subprocess.run([sys.executable, DRIVER, str(d), str(fp), "0", "0:41"], capture_output=True, text=True, env=e, timeout=600)
    assert run.returncode != 0 and "NAN_GUARD_HALT enabled, aborting" in run.stderr
