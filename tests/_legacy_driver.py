"""Subprocess helper of the legacy-ABI tests: one ParakeetSessionSafe (the mirror of rust/parakeet_trt) fed a list of pushes.
usage: _legacy_driver.py <model_dir> <features.npy [128,T]> <use_fp16 0|1> <lo:hi> [<lo:hi> ...]
stdout: one line "rc=<code>" per push (0, or the error code the safe wrapper raised), then "event <kind> <text|message>" per polled event.
stderr: whatever the library prints (PARAKEET_DEBUG_TDT_STEPS trace lines, NAN_GUARD alerts)."""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "trt-asr-engine_b200"))
import numpy as np  # noqa: E402

import binding  # noqa: E402


def main():
    model, fpath, fp16 = sys.argv[1], sys.argv[2], sys.argv[3] == "1"
    f = np.load(fpath)
    s = binding.ParakeetSessionSafe(model, 0, use_fp16=fp16)
    for k, spec in enumerate(sys.argv[4:]):
        lo, hi = (int(x) for x in spec.split(":"))
        seg = np.ascontiguousarray(f[:, lo:hi])
        s.set_debug_context("drv", 1, k, lo)
        try:
            s.push_features(seg, hi - lo)
            print("rc=0")
        except RuntimeError as ex:
            print("rc=" + re.search(r"error code (-?\d+)", str(ex)).group(1))
        while True:
            ev = s.poll_event()
            if ev is None:
                break
            print(f"event {ev.kind} {ev.text if ev.kind != 'error' else ev.message}")
    s.close()


if __name__ == "__main__":
    main()
