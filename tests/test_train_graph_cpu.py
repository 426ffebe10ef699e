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
    i = torch.arange(T)
    s = s.masked_fill(i[None, :] > i[:, None] + delay, float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n, T, 256)


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
    apply = staticmethod(lambda qkv, delay, p=0.0, seed=0: _attn(qkv, delay))


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
