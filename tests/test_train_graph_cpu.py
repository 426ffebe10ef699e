"""Host logic of the training graph (fseend_b200.train_graph / autograd compositions): with the native Functions replaced
by plain torch stand-ins, the composed forward + backward equals the oracle's float64 autograd to rounding.  This pins
the graph wiring (residuals, views, conv-as-GEMM unfolding, split ``convert`` weights, output slicing) on CPU; the
kernels themselves are pinned on the GPU in tests/test_train_ops_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

from oracle import fs_eend_oracle as O


def _attn(qkv, delay):
    n, T, _ = qkv.shape
    q, k, v = (t.reshape(n, T, 4, 64).transpose(1, 2) for t in qkv.split(256, dim=-1))
    s = (q * 0.125) @ k.transpose(-1, -2)
    i = torch.arange(T, device=qkv.device)
    s = s.masked_fill(i[None, :] > i[:, None] + delay, float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n, T, 256)



def _causal4(qkv, delay):
    """[n, T, 768] or the decoder's interleaved [B, T, S, 768] (sequence (b, s) over T)."""
    if qkv.dim() == 3:
        return _attn(qkv, delay)
    B, T, S, _ = qkv.shape
    o = _attn(qkv.transpose(1, 2).reshape(B * S, T, 768), delay)
    return o.reshape(B, S, T, 256).transpose(1, 2)


class _Lin:
    @staticmethod
    def apply(x, w, b, act):
        y = F.linear(x, w, b)
        return torch.relu(y) if act == "relu" else y


class _AddLn:
    @staticmethod
    def apply(x, r, g, b, eps):
        return F.layer_norm(x if r is None else x + r, (256,), g, b, eps)


class _Ffn:
    @staticmethod
    def apply(x, w1, b1, w2, b2):
        return F.linear(torch.relu(F.linear(x, w1, b1)), w2, b2)


class _L2:
    apply = staticmethod(lambda x: x / torch.norm(x, dim=-1, keepdim=True))


class _Head:
    apply = staticmethod(lambda emb, att: (emb[:, :, None, :] * att).sum(-1))


def _bn(bn, x):
    return bn(x.transpose(1, 2)).transpose(1, 2).contiguous()


class _Causal:
    apply = staticmethod(lambda qkv, delay, p=0.0, seed=0: _causal4(qkv, delay))


class _Spk:
    apply = staticmethod(lambda qkv, p=0.0, seed=0: _attn(qkv, 1 << 20))


@pytest.mark.parametrize("mask_delay", [0, 2])
def test_train_graph_wiring_matches_oracle_autograd(monkeypatch, mask_delay):
    import fseend_b200.autograd as A
    import fseend_b200.train_graph as G
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    for mod in (A, G):
        monkeypatch.setattr(mod, "LinearFn", _Lin)
        monkeypatch.setattr(mod, "AddLayerNormFn", _AddLn)
    monkeypatch.setattr(A, "FfnFn", _Ffn)
    monkeypatch.setattr(A, "CausalAttnFn", _Causal)
    monkeypatch.setattr(A, "SpeakerAttnFn", _Spk)
    monkeypatch.setattr(G, "L2NormFn", _L2)
    monkeypatch.setattr(G, "HeadFn", _Head)
    monkeypatch.setattr(G, "batch_norm_forward", _bn)
    monkeypatch.setattr(G, "_require_device", lambda dev: None)
    sd = O.random_state_dict(seed=11, enc_n_layers=1, dec_n_layers=1)
    m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=1, dec_n_layers=1,
                                       dropout=0.0, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048,
                                       mask_delay=mask_delay)
    m.load_state_dict(sd)
    m = m.double().train()
    m.enc.bn.eval()
    lens, n_spks = [60, 41], [3, 2]
    src, _ = O.synthetic_features(2, 60, seed=3, lens=lens)
    g = torch.Generator().manual_seed(9)
    tgt = [(torch.rand(l, n, generator=g) < 0.4).double() for l, n in zip(lens, n_spks)]
    out, el, embs, atts = G.fs_forward_train(m, [s.double() for s in src], tgt, lens)
    loss = G.standard_loss_train(out, tgt, 1) + el
    loss.backward()
    sd64 = {k: v.double().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe"))
            for k, v in sd.items()}
    cfg = O.Cfg(enc_n_layers=1, dec_n_layers=1, mask_delay=mask_delay)
    out_r, el_r, embs_r, atts_r = O.forward(sd64, [s.double() for s in src], tgt, lens, cfg)
    bce = sum(F.binary_cross_entropy_with_logits(y[1:], t[:len(t) - 1]) * (len(y) - 1) for y, t in zip(out_r, tgt))
    loss_r = bce / (sum(lens) - len(lens)) + el_r
    loss_r.backward()
    assert abs(loss.item() - loss_r.item()) < 1e-12
    for a, b in zip(atts, atts_r):
        assert a.shape == b.shape and (a - b).abs().max() < 1e-12
    n_checked = 0
    for name, p in m.named_parameters():
        r = sd64[name].grad
        if r is None:
            assert p.grad is None, name         # dead parameters (dec.encoder*, norm12) on both sides
            continue
        assert (p.grad - r).abs().max() <= 1e-10 * (1 + r.abs().max()), name
        n_checked += 1
    assert n_checked >= 30


TRAIN_CASES = {
    # must match tests/golden/make_golden_train.py
    "train_e2d1": (21, 2, 1, [140, 101, 77], [4, 3, 4], 0, 0),
    "train_e1d2_delay": (22, 1, 2, [90, 64], [3, 5], 2, 1),
}


def load_train_golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "fs_train_grads.json")) as f:
        return json.load(f)


def check_against_reference_grads(m, rec, bce, emb_loss, out, rel):
    """Loss values, logits checksum, BatchNorm running statistics and every parameter gradient (norm + 24 sampled
    elements) against the REAL reference's autograd (tests/golden/fs_train_grads.json)."""
    assert abs(float(bce.detach()) - rec["bce"]) <= rel * abs(rec["bce"])
    assert abs(float(emb_loss.detach()) - rec["emb_loss"]) <= rel * abs(rec["emb_loss"])
    assert abs(float(sum(o.detach().double().sum() for o in out)) - rec["logit_sum"]) <= rel * rec["logit_abs_sum"]
    assert abs(float(m.enc.bn.running_mean.sum()) - rec["running_mean_sum"]) <= rel * (1 + abs(rec["running_mean_sum"]))
    assert abs(float(m.enc.bn.running_var.sum()) - rec["running_var_sum"]) <= rel * (1 + abs(rec["running_var_sum"]))
    n = 0
    for name, p in m.named_parameters():
        g = rec["grads"][name]
        if g is None:
            assert p.grad is None, name
            continue
        flat = p.grad.detach().double().reshape(-1).cpu()
        assert abs(float(flat.norm()) - g["norm"]) <= rel * g["norm"] + 1e-12, name
        got = flat[torch.tensor(g["idx"])]
        want = torch.tensor(g["val"], dtype=torch.float64)
        assert (got - want).abs().max().item() <= rel * g["max_abs"] + 1e-12, name
        n += 1
    assert n >= 40


@pytest.mark.parametrize("case", list(TRAIN_CASES))
def test_train_graph_matches_real_reference_autograd(monkeypatch, case):
    """The train graph's wiring (torch stand-ins for the kernels, float64, BatchNorm in training mode) against gradients
    produced by the real reference model + its own standard_loss (goldens): equal to rounding."""
    import fseend_b200.autograd as A
    import fseend_b200.train_graph as G
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    for mod in (A, G):
        monkeypatch.setattr(mod, "LinearFn", _Lin)
        monkeypatch.setattr(mod, "AddLayerNormFn", _AddLn)
    monkeypatch.setattr(A, "FfnFn", _Ffn)
    monkeypatch.setattr(A, "CausalAttnFn", _Causal)
    monkeypatch.setattr(A, "SpeakerAttnFn", _Spk)
    monkeypatch.setattr(G, "L2NormFn", _L2)
    monkeypatch.setattr(G, "HeadFn", _Head)
    monkeypatch.setattr(G, "batch_norm_forward", _bn)
    monkeypatch.setattr(G, "_require_device", lambda dev: None)
    wseed, ne, nd, lens, n_spks, md, ld = TRAIN_CASES[case]
    sd = O.random_state_dict(seed=wseed, enc_n_layers=ne, dec_n_layers=nd)
    m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=ne, dec_n_layers=nd,
                                       dropout=0.0, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048, mask_delay=md)
    m.load_state_dict(sd)
    m = m.double().train()
    src, _ = O.synthetic_features(len(lens), max(lens), seed=wseed, lens=lens)
    tgt = [t.double() for t in O.synthetic_labels(wseed, lens, n_spks)]
    out, el, _, _ = G.fs_forward_train(m, [s.double() for s in src], tgt, lens)
    bce = G.standard_loss_train(out, tgt, ld)
    (bce + el).backward()
    check_against_reference_grads(m, load_train_golden()[case], bce, el, out, rel=1e-9)
