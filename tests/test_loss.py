"""Training-step label pipeline + standard_loss: the CPU oracle against goldens produced by the REAL
reference code (tests/golden/make_golden_loss.py), and the GPU kernels (through the C ABI and the
train/utils/loss.py drop-in) against both.  Labels are exact (0/1 values, integer re-ordering); the loss is fp32 with a
different summation order: relative tolerance 2e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as L

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_golden.npz"))
CASES = {"loss_B3": (0, [120, 77, 200], [2, 4, 3], 0), "loss_B2_delay": (1, [300, 64], [4, 1], 3),
         "loss_B1": (2, [50], [3], 0)}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_labels_and_loss_match_reference(name):
    seed, lens, n_spks, delay = CASES[name]
    labels, logits = L.synthetic_batch(seed, lens, n_spks)
    tgt = L.prepare_labels(labels)
    for b, t in enumerate(tgt):
        assert np.array_equal(t.numpy(), GOLD[f"{name}_tgt_{b}"])
    assert abs(L.standard_loss(logits, tgt, delay) - float(GOLD[f"{name}_loss"])) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_labels_and_loss_match_reference(name):
    from fseend_b200.loss import prepare_labels, standard_loss
    seed, lens, n_spks, delay = CASES[name]
    labels, logits = L.synthetic_batch(seed, lens, n_spks)
    tgt = prepare_labels([l.cuda() for l in labels])
    for b, t in enumerate(tgt):
        assert np.array_equal(t.cpu().numpy(), GOLD[f"{name}_tgt_{b}"])          # exact
    loss = standard_loss([y.cuda() for y in logits], tgt, label_delay=delay)
    ref = float(GOLD[f"{name}_loss"])
    assert abs(loss.item() - ref) < 2e-6 * max(1.0, abs(ref)) + 2e-6
    again = standard_loss([y.cuda() for y in logits], tgt, label_delay=delay)
    assert loss.item() == again.item()                                            # fixed-order reduction


@pytest.mark.gpu
def test_gpu_loss_at_training_batch_shape():
    """BASELINE configs[1] shape: B=64, T=500, 4 speakers (+2) — kernel vs the fp64 oracle."""
    from fseend_b200.loss import prepare_labels, standard_loss
    labels, logits = L.synthetic_batch(7, [500] * 64, [4] * 64)
    tgt_ref = L.prepare_labels(labels)
    tgt = prepare_labels([l.cuda() for l in labels])
    assert all(np.array_equal(a.cpu().numpy(), b.numpy()) for a, b in zip(tgt, tgt_ref))
    loss = standard_loss([y.cuda() for y in logits], tgt)
    ref = L.standard_loss(logits, tgt_ref)
    assert abs(loss.item() - ref) < 3e-6 * abs(ref)


@pytest.mark.gpu
def test_label_prepare_never_active_speakers_go_last_and_ties_are_stable():
    from fseend_b200 import native as N
    lab = torch.zeros(1, 10, 4)
    lab[0, 3:, 2] = 1          # speaker 2 starts first
    lab[0, 5:, 0] = 1          # speakers 0 and 3 start together: original order kept
    lab[0, 5:, 3] = 1          # speaker 1 never speaks
    out, perm = N.op_label_prepare(lab.cuda())
    assert perm.cpu().tolist() == [[2, 0, 3, 1]]
    assert out.shape == (1, 10, 6) and float(out[0, 0, 0]) == 1.0 and float(out[0, 4, 0]) == 0.0
    assert torch.equal(out[0, :, 1].cpu(), lab[0, :, 2]) and float(out[0, :, 5].abs().sum()) == 0.0
