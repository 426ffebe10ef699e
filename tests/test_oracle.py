"""Pin the CPU oracle (oracle/fs_eend_oracle.py) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import fs_eend_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# must mirror tests/golden/make_golden.py::CASES
CASES = {
    "c1_T500_S4": (0, True, [500], 4, 0),
    "c1_T500_S6": (0, True, [500], 6, 0),
    "ragged_S6": (1, True, [200, 137, 64, 19], 6, 0),
    "default_init_S4": (2, False, [96, 96], 4, 0),
    "maskdelay2_S5": (3, True, [150, 90], 5, 2),
}


# must mirror tests/golden/make_golden.py::WIDE_CASES (logits spanning most of the cosine range)
WIDE_CASES = {
    "wide_pos_S6": (0, 1.5, [300, 211], 6),
    "wide_neg_S6": (1, -1.5, [300, 211], 6),
}


def load_wide_case(name):
    wseed, alpha, lens, S = WIDE_CASES[name]
    sd = O.wide_state_dict(seed=wseed, alpha=alpha, S=S)
    src, lens = O.synthetic_features(len(lens), max(lens), lens=lens)
    return sd, src, lens, S, np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", list(WIDE_CASES))
def test_oracle_matches_reference_on_wide_logits(name):
    sd, src, lens, S, g = load_wide_case(name)
    with torch.no_grad():
        out, _, _ = O.test(sd, src, lens, S, O.Cfg())
    allv = np.concatenate([g[f"logits_{i}"].ravel() for i in range(len(lens))])
    assert allv.max() - allv.min() > 0.7 and allv.std() > 0.2           # the case really is wide
    for i, o in enumerate(out):
        assert np.abs(o.numpy() - g[f"logits_{i}"]).max() < 2e-5


def load_case(name):
    wseed, trained, lens, S, md = CASES[name]
    sd = O.random_state_dict(seed=wseed, trained_like=trained)
    src, lens = O.synthetic_features(len(lens), max(lens), lens=lens)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    return sd, src, lens, S, O.Cfg(mask_delay=md), g


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_logits(name):
    sd, src, lens, S, cfg, g = load_case(name)
    with torch.no_grad():
        out, emb, _ = O.test(sd, src, lens, S, cfg)
    for i, o in enumerate(out):
        ref = g[f"logits_{i}"]
        assert o.shape == ref.shape
        assert np.abs(o.numpy() - ref).max() < 2e-5          # fp32 summation-order noise only
        stride = int(g["emb_stride"][i])
        assert np.abs(emb[i].numpy()[::stride] - g[f"emb_{i}"]).max() < 2e-5


@pytest.mark.parametrize("name", ["ragged_S6", "maskdelay2_S5"])
def test_oracle_forward_and_emb_loss(name):
    sd, src, lens, S, cfg, g = load_case(name)
    gen = torch.Generator().manual_seed(123)
    tgt = [(torch.rand(l, S, generator=gen) > 0.6).float() for l in lens]
    with torch.no_grad():
        out, loss, _, _ = O.forward(sd, src, tgt, lens, cfg)
    assert abs(loss.item() - float(g["emb_consis_loss"])) < 1e-5
    assert np.abs(out[0].numpy() - g["fwd_logits_0"]).max() < 2e-5


def test_oracle_streaming_matches_reference_stream():
    g = np.load(os.path.join(GOLD, "stream_T60_S6.npz"))
    sd = O.random_state_dict(seed=4, trained_like=True)
    src, _ = O.synthetic_features(1, 60)
    with torch.no_grad():
        ys = O.stream_all(sd, src[0][None], 6, O.Cfg())[0]
    assert ys.shape == g["stream"].shape
    assert np.abs(ys.numpy() - g["stream"]).max() < 2e-5
    assert np.abs(ys.numpy() - g["batch"]).max() < 2e-5     # the reference's own invariant: stream == batch


def test_causality_chunk_invariance():
    """Truncating the input leaves earlier logits unchanged up to the conv look-ahead (SURVEY §4 idea)."""
    sd = O.random_state_dict(seed=5)
    src, _ = O.synthetic_features(1, 80)
    cfg = O.Cfg()
    with torch.no_grad():
        full = O.test(sd, src, [80], 4, cfg)[0][0]
        cut = O.test(sd, [src[0][:50]], [50], 4, cfg)[0][0]
    assert (full[:41] - cut[:41]).abs().max() < 1e-5        # frames < 50 - 9 see identical context


def test_fp16_operand_emulation_within_tolerance():
    """The design decision of DESIGN.md §numerics: fp16 operands + fp32 accumulation keep logits within
    1e-3 of the fp32 reference, bf16 does not (measured ~1.5e-3)."""
    sd, src, lens, S, cfg, g = load_case("ragged_S6")
    q16 = lambda x: x.half().float()
    qb = lambda x: x.bfloat16().float()
    with torch.no_grad():
        o16 = O.test(sd, src, lens, S, cfg, quant=q16)[0]
        ob = O.test(sd, src, lens, S, cfg, quant=qb)[0]
    e16 = max(np.abs(o.numpy() - g[f"logits_{i}"]).max() for i, o in enumerate(o16))
    eb = max(np.abs(o.numpy() - g[f"logits_{i}"]).max() for i, o in enumerate(ob))
    assert e16 < 5e-4
    assert eb > e16


def test_sdpa_fast_path_of_the_cpu_arm_is_the_same_arithmetic():
    sd, src, lens, S, cfg, g = load_case("maskdelay2_S5")
    with torch.no_grad():
        a = O.test(sd, src, lens, S, cfg)[0]
        O.USE_SDPA = True
        try:
            b = O.test(sd, src, lens, S, cfg)[0]
        finally:
            O.USE_SDPA = False
    assert max((x - y).abs().max().item() for x, y in zip(a, b)) < 2e-5
