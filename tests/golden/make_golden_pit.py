"""Golden values of the permutation-invariant losses, produced by the REAL reference functions
(/root/reference/FS-EEND/train/utils/loss.py: batch_pit_loss :98-116, batch_pit_n_speaker_loss :257-327,
batch_pit_n_speaker_loss_label_delay :329-403; torchmetrics, absent here and unused by them, is stubbed).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_pit.py
"""
import json
import os
import sys
import types
import warnings

import torch

sys.dont_write_bytecode = True
stub = types.ModuleType("torchmetrics")
stub.PermutationInvariantTraining = object
sys.modules.setdefault("torchmetrics", stub)
sys.path.insert(0, "/root/reference/FS-EEND")
from train.utils.loss import batch_pit_loss, batch_pit_n_speaker_loss, batch_pit_n_speaker_loss_label_delay  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.loss_oracle import synthetic_pit_batch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"pit_B3": (0, [120, 77, 200], [2, 4, 3]), "pit_B2": (1, [300, 64], [3, 1]), "pit_B1": (2, [50], [4])}


def perm_of(orig, permuted):
    """column permutation p with permuted == orig[:, p] (columns of the synthetic labels are distinct)."""
    return [next(j for j in range(orig.shape[1]) if torch.equal(orig[:, j], permuted[:, i])) for i in range(permuted.shape[1])]


def main():
    warnings.filterwarnings("ignore")
    rec = {}
    for name, (seed, lens, n_spks) in CASES.items():
        ys, ts = synthetic_pit_batch(seed, lens, n_spks)
        l1, lab1 = batch_pit_n_speaker_loss([y.clone() for y in ys], [t.clone() for t in ts], n_spks)
        l2, lab2 = batch_pit_n_speaker_loss_label_delay([y.clone() for y in ys], [t.clone() for t in ts], n_spks, 2)
        # batch_pit_loss takes per-recording column counts
        ys_c = [y[:, :n].clone() for y, n in zip(ys, n_spks)]
        ts_c = [t[:, :n].clone() for t, n in zip(ts, n_spks)]
        l3, lab3 = batch_pit_loss(ys_c, ts_c, label_delay=1)
        rec[name] = {"n_speaker_loss": float(l1), "n_speaker_loss_delay2": float(l2), "pit_loss_delay1": float(l3),
                     "label_sums_n_speaker": [[float(c) for c in l.sum(0)] for l in lab1],
                     "label_sums_delay2": [[float(c) for c in l.sum(0)] for l in lab2],
                     "label_sums_pit": [[float(c) for c in l.sum(0)] for l in lab3]}
        print(name, rec[name]["n_speaker_loss"], rec[name]["n_speaker_loss_delay2"], rec[name]["pit_loss_delay1"])
    json.dump(rec, open(os.path.join(HERE, "pit_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
