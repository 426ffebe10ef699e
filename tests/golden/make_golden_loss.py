"""Golden vectors for the training-step label pipeline and standard_loss, produced by the REAL reference code.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_loss.py

* standard_loss is imported from /root/reference/FS-EEND/train/utils/loss.py (torchmetrics, absent here and unused by
  standard_loss, is stubbed before the import).
* The label pipeline is inline in LightningModule.training_step (train/oln_tfm_enc_dec.py:53-76; pytorch_lightning is
  absent): those source lines are read from the reference file AT RUN TIME and executed unmodified on prepared
  variables — nothing of the reference is stored in this repository.
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F  # noqa: F401  (used by the executed reference lines)

sys.dont_write_bytecode = True
stub = types.ModuleType("torchmetrics")
stub.PermutationInvariantTraining = object
sys.modules.setdefault("torchmetrics", stub)
sys.path.insert(0, "/root/reference/FS-EEND")
from train.utils.loss import standard_loss  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.loss_oracle import synthetic_batch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/FS-EEND/train/oln_tfm_enc_dec.py"


def reference_label_pipeline(feats, labels):
    lines = open(REF).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "clip_lengths = [x.shape[0] for x in feats]" in l)
    end = next(i for i in range(start, len(lines)) if "labels = [l[:ilen, :nspk+2]" in lines[i])
    code = textwrap.dedent("\n".join(lines[start:end + 1]))
    env = {"feats": feats, "labels": labels, "torch": torch, "F": F}
    exec(code, env)
    return env["labels"]


CASES = {"loss_B3": (0, [120, 77, 200], [2, 4, 3], 0), "loss_B2_delay": (1, [300, 64], [4, 1], 3),
         "loss_B1": (2, [50], [3], 0)}


def main():
    rec = {}
    for name, (seed, lens, n_spks, delay) in CASES.items():
        labels, logits = synthetic_batch(seed, lens, n_spks)
        feats = [torch.zeros(T, 1) for T in lens]
        tgt = reference_label_pipeline(feats, [l.clone() for l in labels])
        loss = standard_loss(logits, tgt, label_delay=delay)
        for b, t in enumerate(tgt):
            rec[f"{name}_tgt_{b}"] = t.numpy()
        rec[f"{name}_loss"] = np.array(float(loss), dtype=np.float64)
        print(name, "loss", float(loss), [tuple(t.shape) for t in tgt])
    np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **rec)


if __name__ == "__main__":
    main()
