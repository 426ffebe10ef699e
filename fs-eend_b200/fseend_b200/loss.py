"""Drop-in for the part of the reference's FS-EEND/train/utils/loss.py that the FS-EEND training / validation step uses
(train/oln_tfm_enc_dec.py:35,82,122): ``standard_loss(ys, ts, label_delay=0)``, plus the label pipeline that the
reference keeps inline in ``training_step`` (:51-76) as ``prepare_labels``.  Both run on the GPU through the C ABI
(csrc/loss.cu); the list <-> padded-tensor plumbing stays in torch.  Forward values only: backward is SURVEY §8f N1.
"""
import torch
from torch.nn.utils.rnn import pad_sequence


def _device(ts):
    if not torch.cuda.is_available():
        raise RuntimeError("fseend_b200 losses run on a CUDA sm_100 device only")
    return ts[0].device if ts[0].is_cuda else torch.device("cuda")


def standard_loss(ys, ts, label_delay=0):
    """Reference loss.py:119-125.  ys: B-length list of logits (T_b, C_b); ts: B-length list of labels (T_b, C_b)."""
    from fseend_b200.native import op_bce_loss
    dev = _device(ys)
    ymax = max(v.shape[1] for v in ys)
    y = pad_sequence([torch.nn.functional.pad(v.detach().to(device=dev, dtype=torch.float32), (0, ymax - v.shape[1]))
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
    return op_bce_loss(y.contiguous(), t.contiguous(), lens, ncls, label_delay)


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
