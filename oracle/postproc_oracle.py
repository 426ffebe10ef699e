"""TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and the bench CPU leg; never on the product path).

CPU restatement of the reference's RTTM post-processing, FS-EEND/train/utils/make_rttm.py:10-28 (identical file in
LS-EEND; the same threshold + median filter is used by metrics.py:58-60): threshold, scipy-style zero-padded median
filter along time, run-length scan, RTTM line formatting.  Pinned by tests/golden/rttm_*.json, produced by the REAL
reference function (tests/golden/make_golden_rttm.py).
"""
from collections import defaultdict

import numpy as np
import torch


def decide_median(pred: np.ndarray, threshold: float = 0.5, median: int = 11) -> np.ndarray:
    """make_rttm.py:12-15.  (T, C) float -> (T, C) uint8."""
    d = (pred > threshold).astype(np.int64)
    if median > 1:
        half = median // 2
        T = d.shape[0]
        padded = np.concatenate([np.zeros((half, d.shape[1]), np.int64), d, np.zeros((half, d.shape[1]), np.int64)])
        win = np.stack([padded[k:k + T] for k in range(median)], axis=0)     # zero padded windows
        d = np.sort(win, axis=0)[half]                                       # the median itself, not a vote
    return d.astype(np.uint8)


def make_rttm(rec, pred, frame_shift=80, threshold=0.5, median=11, subsampling=10, sampling_rate=8000):
    """make_rttm.py:10-28: dict speaker -> list of RTTM lines."""
    d = decide_median(np.asarray(pred, dtype=np.float32), threshold, median)
    rttm = defaultdict(list)
    fmt = "SPEAKER {:s} 1 {:7.2f} {:7.2f} <NA> <NA> {:s} <NA>"
    for spk in range(d.shape[1]):
        frames = np.concatenate([[0], d[:, spk].astype(np.int64), [0]])
        changes = np.nonzero(np.diff(frames))[0]
        for s, e in zip(changes[::2], changes[1::2]):
            rttm[str(spk)].append(fmt.format(rec, s * frame_shift * subsampling / sampling_rate,
                                             (e - s) * frame_shift * subsampling / sampling_rate, rec + "_" + str(spk)))
    return rttm


def synthetic_posteriors(T: int, C: int, seed: int) -> torch.Tensor:
    """Smooth random speaker activity (random-walk logits + noise) so that segments, flicker and edges all occur."""
    g = torch.Generator().manual_seed(seed)
    x = torch.cumsum(torch.randn(T, C, generator=g) * 0.6, dim=0)
    x = x - x.mean(dim=0, keepdim=True) + torch.randn(T, C, generator=g) * 0.8
    return torch.sigmoid(x).float()
