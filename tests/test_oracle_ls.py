"""Pin the LS-EEND CPU oracle (oracle/ls_eend_oracle.py) against golden vectors produced by the REAL reference
(tests/golden/make_golden_ls.py).  CPU only.  Tolerance 3e-4: the LS-EEND network is ill-conditioned at a few
frames (per-head LayerNorm with eps 1e-6 over near-constant vectors) — the fp32 restatement and the fp32 reference
already differ by up to 8e-5 through summation order alone, and the fp32 and fp64 oracles by 4e-5."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import fs_eend_oracle as FO
from oracle import ls_eend_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LS_CASES = {
    "ls_T1000_ragged_S6": (0, True, [1000, 730], 6),
    "ls_T300_S4": (1, True, [300], 4),
    "ls_T1200_S10": (2, True, [1200], 10),
}


def load_ls_case(name):
    wseed, trained, lens, S = LS_CASES[name]
    sd = O.random_state_dict(seed=wseed, trained_like=trained)
    src, lens = FO.synthetic_features(len(lens), max(lens), lens=lens)
    return sd, src, lens, S, np.load(os.path.join(GOLD, name + ".npz"))


def load_ls_big_case():
    """BASELINE.json configs[2] shape (tests/golden/make_golden_ls.py::big_goldens)."""
    lens = [2000] * 15 + [1711]
    sd = O.random_state_dict(seed=4, trained_like=True)
    src, lens = FO.synthetic_features(len(lens), max(lens), lens=lens)
    return sd, src, lens, 10, np.load(os.path.join(GOLD, "ls_B16_T2000_S10.npz"))


def test_ls_oracle_matches_reference_at_baseline_shape_sample():
    """One recording of the B=16 x T=2000 x S=10 golden through the oracle (the whole batch takes minutes on CPU)."""
    sd, src, lens, S, g = load_ls_big_case()
    with torch.no_grad():
        out, _, _ = O.test(sd, src[15:16], lens[15:16], S, O.Cfg())
    assert out[0].shape == g["logits_15"].shape
    assert np.abs(out[0].numpy() - g["logits_15"]).max() < 3e-4


@pytest.mark.parametrize("name", list(LS_CASES))
def test_ls_oracle_matches_reference(name):
    sd, src, lens, S, g = load_ls_case(name)
    with torch.no_grad():
        out, emb, _ = O.test(sd, src, lens, S, O.Cfg())
    for i, o in enumerate(out):
        assert o.shape == g[f"logits_{i}"].shape
        assert np.abs(o.numpy() - g[f"logits_{i}"]).max() < 3e-4
        stride = int(g["emb_stride"][i])
        assert np.abs(emb[i].numpy()[::stride] - g[f"emb_{i}"]).max() < 3e-4


def ls_forward_loss_case():
    """Inputs of tests/golden/make_golden_ls.py::forward_loss_golden."""
    sd = O.random_state_dict(seed=0, trained_like=True)
    lens, S = [700, 433], 6
    src, lens = FO.synthetic_features(len(lens), max(lens), lens=lens)
    gen = torch.Generator().manual_seed(123)
    tgt = [(torch.rand(l, n, generator=gen) > 0.6).float() for l, n in zip(lens, (S, S - 2))]
    return sd, src, tgt, lens, np.load(os.path.join(GOLD, "ls_forward_loss_S6.npz"))


def test_ls_oracle_forward_and_masked_emb_loss():
    sd, src, tgt, lens, g = ls_forward_loss_case()
    with torch.no_grad():
        out, loss, emb, att = O.forward(sd, src, tgt, lens, O.Cfg())
    assert abs(loss.item() - float(g["emb_consis_loss"])) < 1e-5
    for i in range(2):
        assert np.abs(out[i].numpy() - g[f"fwd_logits_{i}"]).max() < 3e-4
    assert tuple(att[1].shape) == tuple(g["att_shape_1"])


def test_ls_oracle_one_step_matches_reference_stream():
    g = np.load(os.path.join(GOLD, "ls_stream_T48_S4.npz"))
    sd = O.random_state_dict(seed=3)
    src, _ = FO.synthetic_features(1, 48)
    with torch.no_grad():
        ys = O.stream_all(sd, src[0][None], 4, O.Cfg())[0]
    assert ys.shape == g["stream"].shape
    assert np.abs(ys.numpy() - g["stream"]).max() < 1e-4


def test_ls_state_dict_abi_and_strict_load():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "fs-eend_b200"))
    from nnet.model.onl_conformer_retention_enc_1dcnn_tfm_retention_enc_linear_non_autoreg_pos_enc_l2norm_emb_loss_mask import (
        OnlineConformerRetentionDADiarization)
    m = OnlineConformerRetentionDADiarization(
        n_speakers=8, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1,
        max_seqlen=1000, recurrent_chunk_size=500, feed_forward_expansion_factor=4, dec_dim_feedforward=2048,
        conv_kernel_size=16)
    want = {}
    with open(os.path.join(GOLD, "ls_state_dict_abi.txt")) as f:
        for line in f:
            k, shape, dt = re.match(r"(\S+) (\(.*\)) (\S+)", line.strip()).groups()
            want[k] = (eval(shape), dt)
    got = {k: (tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in m.state_dict().items()}
    assert got == want
    assert sum(p.numel() for p in m.parameters()) == 11_183_616
    m.load_state_dict(O.random_state_dict(0), strict=True)
