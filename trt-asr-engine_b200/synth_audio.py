"""Seeded synthetic 16 kHz audio: white noise + 5 slowly amplitude-modulated sinusoids
between 200 and 3400 Hz, f32 in [-1, 1].  Used by tests and bench.py (there is no network for datasets)."""
from __future__ import annotations

import numpy as np


def synth_clip(seconds: float, seed: int, sr: int = 16000) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    x = 0.05 * rng.standard_normal(n)
    for _ in range(5):
        f = rng.uniform(200.0, 3400.0)
        am_f = rng.uniform(0.5, 4.0)
        ph, am_ph = rng.uniform(0, 2 * np.pi, size=2)
        x += 0.08 * (0.5 + 0.5 * np.sin(2 * np.pi * am_f * t + am_ph)) * np.sin(2 * np.pi * f * t + ph)
    return np.clip(x, -1.0, 1.0).astype(np.float32)
