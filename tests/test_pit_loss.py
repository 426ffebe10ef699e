"""Permutation-invariant losses (reference train/utils/loss.py:98-116, :257-327, :329-403): the CPU oracle and the GPU
path (pair-cost kernel through the C ABI + host permutation search, fseend_b200.loss) against values produced by the REAL
reference functions (tests/golden/make_golden_pit.py)."""
import json
import os

import pytest
import torch

from oracle import loss_oracle as L

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pit_golden.json")))
CASES = {"pit_B3": (0, [120, 77, 200], [2, 4, 3]), "pit_B2": (1, [300, 64], [3, 1]), "pit_B1": (2, [50], [4])}


def sums(labels):
    return [[float(c) for c in l.sum(0)] for l in labels]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference(name):
    seed, lens, n_spks = CASES[name]
    ys, ts = L.synthetic_pit_batch(seed, lens, n_spks)
    g = GOLD[name]
    l1, lab1 = L.batch_pit_n_speaker_loss(ys, ts, n_spks)
    l2, lab2 = L.batch_pit_n_speaker_loss(ys, ts, n_spks, label_delay=2)
    l3, lab3 = L.batch_pit_loss([y[:, :n] for y, n in zip(ys, n_spks)], [t[:, :n] for t, n in zip(ts, n_spks)], 1)
    assert abs(l1 - g["n_speaker_loss"]) < 2e-6 and sums(lab1) == g["label_sums_n_speaker"]
    assert abs(l2 - g["n_speaker_loss_delay2"]) < 2e-6 and sums(lab2) == g["label_sums_delay2"]
    assert abs(l3 - g["pit_loss_delay1"]) < 2e-6 and sums(lab3) == g["label_sums_pit"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_pit_losses_match_reference(name):
    from fseend_b200.loss import batch_pit_loss, batch_pit_n_speaker_loss, batch_pit_n_speaker_loss_label_delay
    seed, lens, n_spks = CASES[name]
    ys, ts = L.synthetic_pit_batch(seed, lens, n_spks)
    ys, ts = [y.cuda() for y in ys], [t.cuda() for t in ts]
    g = GOLD[name]
    l1, lab1 = batch_pit_n_speaker_loss(ys, ts, n_spks)
    l2, lab2 = batch_pit_n_speaker_loss_label_delay(ys, ts, n_spks, 2)
    l3, lab3 = batch_pit_loss([y[:, :n] for y, n in zip(ys, n_spks)], [t[:, :n] for t, n in zip(ts, n_spks)], 1)
    assert abs(l1.item() - g["n_speaker_loss"]) < 3e-6 and sums(lab1) == g["label_sums_n_speaker"]
    assert abs(l2.item() - g["n_speaker_loss_delay2"]) < 3e-6 and sums(lab2) == g["label_sums_delay2"]
    assert abs(l3.item() - g["pit_loss_delay1"]) < 3e-6 and sums(lab3) == g["label_sums_pit"]
    assert all(l.shape == (T, n) for l, T, n in zip(lab1, lens, n_spks))


@pytest.mark.gpu
def test_gpu_pit_pair_costs_kernel_vs_fp64():
    from fseend_b200 import native as N
    ys, ts = L.synthetic_pit_batch(5, [257, 40], [4, 4])
    y = torch.nn.utils.rnn.pad_sequence(ys, batch_first=True).cuda().contiguous()
    t = torch.nn.utils.rnn.pad_sequence(ts, batch_first=True).cuda().contiguous()
    lens = torch.tensor([257, 40], dtype=torch.int32, device="cuda")
    cost = N.op_pit_costs(y, t, lens, label_delay=3).cpu()
    for b in range(2):
        ref = L.pit_pair_costs(ys[b], ts[b], 3)
        assert (cost[b] - ref).abs().max().item() < 1e-4 * ref.abs().max().item()
    again = N.op_pit_costs(y, t, lens, label_delay=3).cpu()
    assert torch.equal(cost, again)                       # fixed-order reduction
