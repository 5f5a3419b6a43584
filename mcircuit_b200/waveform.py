"""Waveform export: samples from ``B200Engine.capture`` -> VCD text for one lane."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def lane_values(samples: np.ndarray, widths: Sequence[int], lane: int = 0):
    """samples: uint64[n_samples, n_bits, words] (bit planes, pins concatenated).
    Returns a list of per-pin integer traces for ``lane``."""
    word, bit = divmod(lane, 64)
    bits = ((samples[:, :, word] >> np.uint64(bit)) & np.uint64(1)).astype(object)
    out, k = [], 0
    for w in widths:
        v = np.zeros(samples.shape[0], dtype=object)
        for b in range(w):
            v += bits[:, k + b] << b
        out.append([int(x) for x in v])
        k += w
    return out


def _ident(i: int) -> str:
    chars = [chr(c) for c in range(33, 127)]
    s = ""
    while True:
        s = chars[i % len(chars)] + s
        i = i // len(chars) - 1
        if i < 0:
            return s


def to_vcd(names: Sequence[str], widths: Sequence[int], samples: np.ndarray, lane: int = 0,
           timescale: str = "1 ns", step_time: int = 1) -> str:
    """Value-change dump of ``lane``: one timestep per sample, only changes are written."""
    traces = lane_values(samples, widths, lane)
    lines = ["$version mcircuit_b200 $end", "$timescale %s $end" % timescale, "$scope module top $end"]
    ids = [_ident(i) for i in range(len(names))]
    for name, w, i in zip(names, widths, ids):
        lines.append("$var wire %d %s %s $end" % (w, i, name.strip("/").replace("/", ".")))
    lines += ["$upscope $end", "$enddefinitions $end"]
    last = [None] * len(names)
    for t in range(samples.shape[0]):
        changes = []
        for k, (w, i) in enumerate(zip(widths, ids)):
            v = traces[k][t]
            if v != last[k]:
                changes.append(("%d%s" % (v, i)) if w == 1 else ("b%s %s" % (format(v, "b"), i)))
                last[k] = v
        if changes or t == 0:
            lines.append("#%d" % (t * step_time))
            lines += changes
    lines.append("#%d" % (samples.shape[0] * step_time))
    return "\n".join(lines) + "\n"
