"""Gradient goldens from the REAL reference's autograd (imported read-only from /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_train.py

For each case the real ``OnlineTransformerDADiarization`` (FS-EEND/nnet/model/onl_tfm_...l2norm.py) is put in
``train()`` mode (dropout 0, BatchNorm with batch statistics over the -1-padded batch), run through
``model(src, tgt, ilens)``; the training loss of ``train/oln_tfm_enc_dec.py:78-85`` — the real ``standard_loss`` of
``train/utils/loss.py`` plus the embedding loss — is back-propagated by torch autograd, in float64 (the reference
arithmetic with rounding removed: gradients of the early layers are ill-conditioned in float32).
Stored per parameter (a few KB in total): L2 norm, sum and 24 sampled elements (fixed seeded indices) of the gradient,
plus the loss values, the logits' checksum and BatchNorm's updated running statistics.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = "/root/reference/FS-EEND"
sys.path.insert(0, REF)

import types  # noqa: E402

stub = types.ModuleType("torchmetrics")          # absent here and unused by standard_loss
stub.PermutationInvariantTraining = object
sys.modules.setdefault("torchmetrics", stub)

from oracle import fs_eend_oracle as O  # noqa: E402

from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization  # noqa: E402
from train.utils.loss import standard_loss  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (weight seed, enc layers, dec layers, lens, n_spks per item (label columns), mask_delay, label_delay)
    "train_e2d1": (21, 2, 1, [140, 101, 77], [4, 3, 4], 0, 0),
    "train_e1d2_delay": (22, 1, 2, [90, 64], [3, 5], 2, 1),
}


def sample_index(name, numel, k=24):
    g = torch.Generator().manual_seed(abs(hash_name(name)) % (2 ** 31))
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def hash_name(name):
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000000007
    return h


def inputs(seed, lens, n_spks):
    src, _ = O.synthetic_features(len(lens), max(lens), seed=seed, lens=lens)
    return src, O.synthetic_labels(seed, lens, n_spks)


def main():
    out = {}
    for name, (wseed, ne, nd, lens, n_spks, md, ld) in CASES.items():
        sd = O.random_state_dict(seed=wseed, enc_n_layers=ne, dec_n_layers=nd)
        m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=ne, dec_n_layers=nd,
                                           dropout=0.0, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048, mask_delay=md)
        m.load_state_dict(sd, strict=True)
        m = m.double().train()
        # The reference builds its causal mask in float32 (model file :152-155).  Run in float64, torch mis-applies a float
        # mask whose dtype differs from the activations' (the float64 encoder output is 5.6 away from the float32 run), so
        # the maker casts the reference's own mask to float64; everything else is the reference's code.
        for sub in (m.enc, m.dec):
            gen = sub._generate_square_subsequent_mask
            sub._generate_square_subsequent_mask = (lambda g: (lambda sz, device: g(sz, device).double()))(gen)
        src, tgt = inputs(wseed, lens, n_spks)
        preds, emb_loss, _, _ = m([s.double() for s in src], [t.double() for t in tgt], lens)
        bce = standard_loss(preds, [t.double() for t in tgt], label_delay=ld)
        (bce + emb_loss).backward()
        rec = {"bce": float(bce), "emb_loss": float(emb_loss), "logit_sum": float(sum(p.sum() for p in preds)),
               "logit_abs_sum": float(sum(p.abs().sum() for p in preds)),
               "running_mean_sum": float(m.enc.bn.running_mean.sum()), "running_var_sum": float(m.enc.bn.running_var.sum()),
               "grads": {}}
        for pname, p in m.named_parameters():
            if p.grad is None:
                rec["grads"][pname] = None
                continue
            gflat = p.grad.reshape(-1)
            idx = sample_index(pname, gflat.numel())
            rec["grads"][pname] = {"norm": float(gflat.norm()), "sum": float(gflat.sum()), "max_abs": float(gflat.abs().max()),
                                   "idx": idx.tolist(), "val": [float(v) for v in gflat[idx]]}
        out[name] = rec
        print(name, "bce", rec["bce"], "emb", rec["emb_loss"], "params", len(rec["grads"]))
    with open(os.path.join(HERE, "fs_train_grads.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "fs_train_grads.json"), os.path.getsize(os.path.join(HERE, "fs_train_grads.json")), "bytes")


if __name__ == "__main__":
    main()
