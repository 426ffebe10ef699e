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


# ----------------------------------------------------------------------------- permutation-invariant losses
def _bce(x, t):
    return torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))


def pit_pair_costs(y, t, label_delay=0):
    """cost[i][j] = sum_t BCE(y[t + delay, i], t[t, j]) in fp64 (the quantity every PIT variant permutes)."""
    yy, tt = y[label_delay:].double(), t[:len(t) - label_delay].double()
    return torch.stack([torch.stack([_bce(yy[:, i], tt[:, j]).sum() for j in range(t.shape[1])])
                        for i in range(y.shape[1])])


def batch_pit_loss(ys, ts, label_delay=0):
    """loss.py:98-116 with pit_loss :69-96: per recording the minimum over label permutations of the mean BCE, times the
    number of frames; summed and divided by the total number of frames."""
    from itertools import permutations
    total, labels = 0.0, []
    for y, t in zip(ys, ts):
        C = t.shape[1]
        cost = pit_pair_costs(y, t, label_delay)
        best, best_p = None, None
        for p in permutations(range(C)):
            v = sum(cost[i, p[i]].item() for i in range(C)) / C
            if best is None or v < best:
                best, best_p = v, p
        total += best
        labels.append(t[..., list(best_p)])
    return total / sum(t.shape[0] for t in ts), labels


def batch_pit_n_speaker_loss(ys, ts, n_speakers_list, label_delay=None):
    """loss.py:257-327 (label_delay None: shorter recordings enter with their -1-padded frames, BCE(-1, -1) each) and
    :329-403 (label_delay given: padded frames excluded)."""
    from itertools import permutations
    C = max(n_speakers_list)
    Tmax = max(t.shape[0] for t in ts)
    pad_bce = _bce(torch.tensor(-1.0, dtype=torch.float64), torch.tensor(-1.0, dtype=torch.float64)).item()
    total, labels = 0.0, []
    for y, t, n in zip(ys, ts, n_speakers_list):
        cost = pit_pair_costs(y, t, 0 if label_delay is None else label_delay)
        if label_delay is None:
            cost = cost + (Tmax - t.shape[0]) * pad_bce
        best, best_p = None, None
        for p in permutations(range(C)):
            if list(p[n:]) != sorted(p[n:]):
                continue
            v = sum(cost[i, p[i]].item() for i in range(C)) / C
            if best is None or v < best:
                best, best_p = v, p
        total += best
        labels.append(t[:, list(best_p)][:, :n])
    return total / sum(t.shape[0] for t in ts), labels


def synthetic_pit_batch(seed, lens, n_spks):
    """Labels (T_b, n_b) 0/1 and logits correlated with a random permutation of them (so that the best permutation is not
    the identity), both padded to C = max(n_spks) columns as the reference's callers do (pad_labels / pad_preds)."""
    g = torch.Generator().manual_seed(seed)
    C = max(n_spks)
    ys, ts = [], []
    for T, n in zip(lens, n_spks):
        lab = (torch.rand(T, n, generator=g) > 0.5).float()
        perm = torch.randperm(n, generator=g)
        y = (lab[:, perm] * 2 - 1) * 1.5 + torch.randn(T, n, generator=g)
        ts.append(torch.nn.functional.pad(lab, (0, C - n)))
        ys.append(torch.nn.functional.pad(y, (0, C - n), value=-4.0))
    return ys, ts
