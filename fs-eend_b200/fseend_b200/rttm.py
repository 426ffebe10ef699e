"""Drop-in for the reference's FS-EEND/train/utils/make_rttm.py (same path in LS-EEND): posteriors -> RTTM lines.

``make_rttm(rec, pred, frame_shift=80, threshold=0.5, median=11, subsampling=10, sampling_rate=8000)`` returns the same
``defaultdict(list)`` keyed by ``str(speaker index)`` with byte-identical ``SPEAKER ...`` lines (reference :10-28; callers:
streaming_infer_dia.py:99-103, dia_pred.py:58, train/oln_tfm_enc_dec.py:263).  The threshold + median filter run on the
GPU (csrc/elementwise.cu: decide_median_kernel, through the C ABI); only the run-length scan over the tiny
(T, n_spk) decision matrix and the string formatting stay on the host.  No CPU fallback for the filter.
"""
from collections import defaultdict

import torch

FMT = "SPEAKER {:s} 1 {:7.2f} {:7.2f} <NA> <NA> {:s} <NA>"


def decisions(pred, threshold=0.5, median=11):
    """(T, n_spk) posteriors (any device) -> uint8 CUDA tensor of filtered 0/1 decisions."""
    from fseend_b200.native import op_decide_median
    if not torch.cuda.is_available():
        raise RuntimeError("fseend_b200 post-processing runs on a CUDA sm_100 device only")
    p = torch.as_tensor(pred).detach().to(device="cuda", dtype=torch.float32).contiguous()
    return op_decide_median(p, threshold, median)


def segments(dec):
    """uint8 (T, n_spk) decisions -> list per speaker of (start_frame, end_frame) with end exclusive."""
    d = dec.to("cpu").numpy()
    out = []
    for spk in range(d.shape[1]):
        col = d[:, spk]
        segs, start = [], None
        for t in range(len(col)):
            if col[t] and start is None:
                start = t
            elif not col[t] and start is not None:
                segs.append((start, t))
                start = None
        if start is not None:
            segs.append((start, len(col)))
        out.append(segs)
    return out


def make_rttm(rec, pred, frame_shift=80, threshold=0.5, median=11, subsampling=10, sampling_rate=8000):
    rttm = defaultdict(list)
    dec = decisions(pred, threshold, median)
    for spkid, segs in enumerate(segments(dec)):
        for s, e in segs:
            rttm[str(spkid)].append(FMT.format(
                rec,
                s * frame_shift * subsampling / sampling_rate,
                (e - s) * frame_shift * subsampling / sampling_rate,
                rec + "_" + str(spkid)))
    return rttm
