"""Drop-in for the part of the reference's FS-EEND/train/utils/loss.py that the FS-EEND training / validation step uses
(train/oln_tfm_enc_dec.py:35,82,122): ``standard_loss(ys, ts, label_delay=0)``, plus the label pipeline that the
reference keeps inline in ``training_step`` (:51-76) as ``prepare_labels``, plus the permutation-invariant losses of the
fine-tuning / offline harnesses (``batch_pit_loss`` :98-116, ``batch_pit_n_speaker_loss`` :257-327 and its label-delay
form :329-403).  The frame sums run on the GPU through the C ABI (csrc/loss.cu); the list <-> padded-tensor plumbing
and the search over the C! permutations of a C x C cost matrix stay on the host.  ``standard_loss`` is differentiable (native value, closed-form
gradient); the PIT losses return values only.
"""
from itertools import permutations


import torch
from torch.nn.utils.rnn import pad_sequence


def _device(ts):
    if not torch.cuda.is_available():
        raise RuntimeError("fseend_b200 losses run on a CUDA sm_100 device only")
    return ts[0].device if ts[0].is_cuda else torch.device("cuda")


class _BceLossFn(torch.autograd.Function):
    """standard_loss value from the native kernel (csrc/loss.cu); gradient (sigmoid(y) - label) / (classes * frames) on
    the frames / classes the loss covers, as the reference's autograd produces for loss.py:119-125."""

    @staticmethod
    def forward(ctx, y, t, lens, ncls, label_delay):
        from fseend_b200.native import op_bce_loss
        ctx.save_for_backward(y, t, lens, ncls)
        ctx.delay = int(label_delay)
        return op_bce_loss(y.detach().contiguous(), t.contiguous(), lens, ncls, label_delay)

    @staticmethod
    def backward(ctx, g):
        y, t, lens, ncls = ctx.saved_tensors
        B, T, Cn = y.shape
        d = ctx.delay
        tt = torch.zeros_like(y)
        tt[:, d:, :t.shape[2]] = t[:, :T - d]
        fr = torch.arange(T, device=y.device)[None, :, None]
        valid = (fr >= d) & (fr < lens[:, None, None]) & (torch.arange(Cn, device=y.device)[None, None, :] < ncls[:, None, None])
        n_frames = (lens.sum() - d * B).to(torch.float32)
        grad = (torch.sigmoid(y) - tt) * valid / (ncls[:, None, None].to(torch.float32) * n_frames)
        return grad * g, None, None, None, None


def standard_loss(ys, ts, label_delay=0):
    """Reference loss.py:119-125.  ys: B-length list of logits (T_b, C_b); ts: B-length list of labels (T_b, C_b).
    Differentiable w.r.t. ys when they carry autograd history (the training step, train/oln_tfm_enc_dec.py:82)."""
    dev = _device(ys)
    ymax = max(v.shape[1] for v in ys)
    y = pad_sequence([torch.nn.functional.pad(v.to(device=dev, dtype=torch.float32), (0, ymax - v.shape[1]))
                      for v in ys], batch_first=True)
    cmax = max(t.shape[1] for t in ts)
    t = pad_sequence([torch.nn.functional.pad(v.detach().to(device=dev, dtype=torch.float32), (0, cmax - v.shape[1]))
                      for v in ts], batch_first=True)
    if y.shape[1] != t.shape[1]:
        raise ValueError("predictions and labels must cover the same frames")
    lens = torch.tensor([v.shape[0] for v in ts], device=dev, dtype=torch.int32)
    ncls = torch.tensor([v.shape[1] for v in ts], device=dev, dtype=torch.int32)
    for yy, tt in zip(ys, ts):
        if tuple(yy.shape) != tuple(tt.shape):
            # the reference's binary_cross_entropy_with_logits raises on any shape mismatch (loss.py:121-123)
            raise ValueError("each prediction must have exactly its label's shape (frames, classes)")
    if t.shape[2] != y.shape[2]:
        t = torch.nn.functional.pad(t, (0, y.shape[2] - t.shape[2]))
    return _BceLossFn.apply(y, t, lens, ncls, label_delay)


def prepare_labels(labels, clip_lengths=None):
    """Reference train/oln_tfm_enc_dec.py:53-76: list of (T_b, n_spk_b) 0/1 activity matrices -> list of
    (T_b, n_spk_b + 2) training targets: silence | speakers ordered by first appearance | "no speaker" zeros."""
    from fseend_b200.native import op_label_prepare
    dev = _device(labels)
    n_spks = [l.shape[1] for l in labels]
    max_spk = max(n_spks)
    lens = [l.shape[0] for l in labels] if clip_lengths is None else list(clip_lengths)
    lab = pad_sequence([torch.nn.functional.pad(l.detach().to(device=dev, dtype=torch.float32), (0, max_spk - l.shape[1]))
                        for l in labels], batch_first=True).contiguous()
    out, _ = op_label_prepare(lab)
    return [o[:ilen, :n + 2] for o, ilen, n in zip(out, lens, n_spks)]


def _pit_costs(ys, ts, label_delay, pad_term, C):
    from fseend_b200.native import op_pit_costs
    dev = _device(ys)
    y = pad_sequence([torch.nn.functional.pad(v.detach().to(device=dev, dtype=torch.float32), (0, C - v.shape[1]))
                      for v in ys], batch_first=True).contiguous()
    t = pad_sequence([torch.nn.functional.pad(v.detach().to(device=dev, dtype=torch.float32), (0, C - v.shape[1]))
                      for v in ts], batch_first=True).contiguous()
    lens = torch.tensor([v.shape[0] for v in ts], device=dev, dtype=torch.int32)
    return op_pit_costs(y, t, lens, label_delay, pad_term).cpu()


def batch_pit_loss(ys, ts, label_delay=0):
    """Reference loss.py:98-116 (pit_loss :69-96 per recording).  ys, ts: B-length lists of (T_b, C_b) logits / labels.
    Returns (loss, list of permuted labels)."""
    for y, t in zip(ys, ts):
        if tuple(y.shape) != tuple(t.shape):
            raise ValueError("each prediction must have exactly its label's shape (frames, classes)")
    C = max(t.shape[1] for t in ts)
    cost = _pit_costs(ys, ts, label_delay, False, C)
    total, labels = 0.0, []
    for b, t in enumerate(ts):
        cb = t.shape[1]
        best, best_p = None, None
        for p in permutations(range(cb)):                       # same enumeration order as the reference: first minimum
            v = sum(float(cost[b, i, p[i]]) for i in range(cb)) / cb
            if best is None or v < best:
                best, best_p = v, p
        total += best
        labels.append(t[..., list(best_p)])
    n_frames = sum(t.shape[0] for t in ts)
    return torch.tensor(total / n_frames, dtype=torch.float32, device=_device(ys)), labels


def batch_pit_n_speaker_loss(ys, ts, n_speakers_list, label_delay=None):
    """Reference loss.py:257-327 (label_delay None) and :329-403 (label_delay given: frames beyond each recording's length
    are excluded instead of entering as -1-padded frames).  ys, ts: B-length lists of (T_b, C) with the same C =
    max(n_speakers_list) columns (pad_labels / pad_preds upstream, as in the reference's callers).
    Returns (loss, list of permuted labels cut to n_speakers columns)."""
    C = max(n_speakers_list)
    for y, t in zip(ys, ts):
        if y.shape[1] != C or t.shape[1] != C or y.shape[0] != t.shape[0]:
            raise ValueError("every prediction / label must be (T_b, max(n_speakers_list))")
    cost = _pit_costs(ys, ts, 0 if label_delay is None else label_delay, label_delay is None, C)
    perms = list(permutations(range(C)))
    total, labels = 0.0, []
    for b, (t, n) in enumerate(zip(ts, n_speakers_list)):
        # admissible permutations: every ordering of the n real speakers, the remaining columns kept in ascending order
        # (the reference's select_perm_indices: first permutation, in enumeration order, with that prefix)
        best, best_p = None, None
        for p in perms:
            if list(p[n:]) != sorted(p[n:]):
                continue
            v = float(torch.tensor([float(cost[b, i, p[i]]) for i in range(C)], dtype=torch.float32).mean())
            if best is None or v < best:
                best, best_p = v, p
        total += best
        labels.append(t[:, list(best_p)][:, :n])
    n_frames = sum(t.shape[0] for t in ts)
    return torch.tensor(total / n_frames, dtype=torch.float32, device=_device(ys)), labels


def batch_pit_n_speaker_loss_label_delay(ys, ts, n_speakers_list, label_delay=0):
    return batch_pit_n_speaker_loss(ys, ts, n_speakers_list, label_delay=label_delay)
