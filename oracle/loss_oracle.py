"""TEST INFRASTRUCTURE ONLY (checker for tests/; never on the product path).

CPU restatement of the FS-EEND training-step label pipeline (FS-EEND/train/oln_tfm_enc_dec.py:53-76) and of
standard_loss (FS-EEND/train/utils/loss.py:119-125).  Pinned by tests/golden/loss_golden.npz, produced by running the
REAL reference code (tests/golden/make_golden_loss.py).
"""
import numpy as np
import torch


def prepare_labels(labels, clip_lengths=None):
    """oln_tfm_enc_dec.py:53-76.  labels: list of (T_b, n_spk_b) float 0/1 tensors."""
    n_spks = [l.shape[1] for l in labels]
    max_spk = max(n_spks)
    lens = [l.shape[0] for l in labels] if clip_lengths is None else list(clip_lengths)
    B, T = len(labels), max(l.shape[0] for l in labels)
    lab = torch.zeros(B, T, max_spk)
    for b, l in enumerate(labels):
        lab[b, :l.shape[0], :l.shape[1]] = l
    out = []
    for b in range(B):
        first = []
        for c in range(max_spk):                      # :61-63: (frame index + 1) of the first active frame, inf if none
            nz = torch.nonzero(lab[b, :, c])
            first.append(float(nz[0, 0] + 1) if len(nz) else float("inf"))
        order = sorted(range(max_spk), key=lambda c: (first[c], c))   # :66 argsort (stable here)
        spk = lab[b][:, order]                                          # :67
        silence = 1.0 - lab[b].max(dim=-1)[0]                           # :69
        full = torch.cat([silence[:, None], spk, torch.zeros(T, 1)], dim=-1)   # :70-73
        out.append(full[:lens[b], :n_spks[b] + 2])                      # :75
    return out


def standard_loss(ys, ts, label_delay=0):
    """loss.py:119-125 in fp64."""
    tot = 0.0
    for y, t in zip(ys, ts):
        yy = y[label_delay:, :t.shape[1]].double()
        tt = t[:len(t) - label_delay].double()
        bce = torch.clamp(yy, min=0) - yy * tt + torch.log1p(torch.exp(-yy.abs()))
        tot += bce.mean().item() * (len(y) - label_delay)
    n_frames = sum(t.shape[0] for t in ts) - label_delay * len(ts)
    return tot / n_frames


def synthetic_batch(seed, lens, n_spks):
    """Random activity with delayed speaker onsets (distinct first-appearance frames) + random logits."""
    g = torch.Generator().manual_seed(seed)
    labels, logits = [], []
    for T, n in zip(lens, n_spks):
        onset = torch.randperm(max(T // 2, n + 1), generator=g)[:n] + 1        # distinct onsets
        act = (torch.rand(T, n, generator=g) > 0.55).float()
        act[torch.arange(T)[:, None] < onset[None, :]] = 0
        for c in range(n):                                                      # make the onset frame itself active
            if int(onset[c]) < T:
                act[int(onset[c]), c] = 1.0
        labels.append(act)
        logits.append(torch.randn(T, n + 2, generator=g) * 2.0)
    return labels, logits
