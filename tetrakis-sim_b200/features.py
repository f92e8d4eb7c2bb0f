"""Spectral / time-series features of a probe series (reference ``tetrakis_sim/evals/features.py``).

Same keys and definitions as the reference so that ``tetrakis-eval generate`` records keep schema
version 1 (``evals/dataset.py:14,146-163``): non-negative-frequency half of the spectrum,
dominant frequency, centroid, bandwidth, peak amplitude, low/high split, the five largest peaks
(bin index + magnitude relative to the maximum), rms and peak of the series.
"""

from __future__ import annotations

import numpy as np

TOPK = 5
EPS = 1e-12


def spectral_features(freq, amp) -> dict:
    f = np.asarray(freq, dtype=float)
    a = np.asarray(amp, dtype=float)
    if f.shape != a.shape:
        raise ValueError("freq and amp must have the same shape")
    keep = f >= 0  # features.py:18-21
    f, a = f[keep], a[keep]
    total = float(a.sum())
    names = ["dominant_freq", "spectral_centroid", "spectral_bandwidth", "peak_amplitude", "low_high_ratio"]
    out = dict.fromkeys(names, 0.0)
    for k in range(TOPK):
        out[f"peak_i_{k}"] = -1.0
        out[f"peak_mag_{k}"] = 0.0
    if a.size == 0 or total <= 0.0:
        return out
    centroid = float((f * a).sum() / total)
    half = int(a.size // 2)
    low = float(a[:half].sum()) if half > 0 else 0.0
    amax = float(a.max())
    out.update(
        dominant_freq=float(f[int(np.argmax(a))]),
        spectral_centroid=centroid,
        spectral_bandwidth=float(np.sqrt(((f - centroid) ** 2 * a).sum() / total)),
        peak_amplitude=amax,
        low_high_ratio=float(low / (float(a[half:].sum()) + EPS)),
    )
    k = min(TOPK, int(a.size))
    top = np.argpartition(a, -k)[-k:]  # features.py:51-52: same selection + ordering calls
    top = top[np.argsort(a[top])[::-1]]
    for j, i in enumerate(top):
        out[f"peak_i_{j}"] = float(int(i))
        out[f"peak_mag_{j}"] = float(a[int(i)] / (amax + EPS))
    return out


def timeseries_features(values) -> dict:
    v = np.asarray(values, dtype=float)
    if v.size == 0:
        return {"ts_rms": 0.0, "ts_peak": 0.0}
    return {"ts_rms": float(np.sqrt(np.mean(v**2))), "ts_peak": float(np.max(np.abs(v)))}
