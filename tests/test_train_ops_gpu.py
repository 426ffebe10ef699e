"""Training building blocks (started): forward + backward of the native Linear(+ReLU), residual +
LayerNorm and causal attention kernels against torch autograd on the same seeded inputs, and the composed post-norm
encoder layer against nn.TransformerEncoderLayer (the class the reference instantiates at FS:model:147).

Tolerance: every tensor is compared as max|native - ref| <= 2e-5 * max|ref| + 1e-6 with the reference computed in
float64 — the split-precision products carry ~22 operand mantissa bits, the fp32 reductions ~1e-6."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

REL, ABS = 2e-5, 1e-6


def close(a, b, what, rel=REL):
    a, b = a.double().cpu(), b.double().cpu()
    err = (a - b).abs().max().item()
    bound = rel * b.abs().max().item() + ABS
    assert err <= bound, f"{what}: max err {err:.3e} > {bound:.3e}"


@pytest.mark.parametrize("rows,K,N,act", [(1000, 256, 2048, "relu"), (777, 2048, 256, "none"), (500, 345, 256, "none"),
                                          (64, 256, 768, "none"), (3, 256, 256, "relu")])
def test_linear_fwd_bwd(built_lib, rows, K, N, act):
    from fseend_b200.autograd import LinearFn
    g = torch.Generator().manual_seed(rows + K)
    x = torch.randn(rows, K, generator=g).cuda().requires_grad_()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().requires_grad_()
    b = (0.1 * torch.randn(N, generator=g)).cuda().requires_grad_()
    dy = torch.randn(rows, N, generator=g).cuda()
    y = LinearFn.apply(x, w, b, act)
    y.backward(dy)
    xr, wr, br = (t.detach().double().requires_grad_() for t in (x, w, b))
    yr = xr @ wr.t() + br
    if act == "relu":
        # the ReLU gate is decided by the native output's sign: rows where the fp64 pre-activation is within rounding
        # of zero would differ by a whole element, so the reference gates with the same mask
        mask = (y.detach() > 0).double()
        yr = yr * mask
    yr.backward(dy.double())
    close(y, yr, "y")
    close(x.grad, xr.grad, "dx")
    close(w.grad, wr.grad, "dw")
    close(b.grad, br.grad, "db")


@pytest.mark.parametrize("rows", [1000, 130])
def test_ffn_fwd_bwd(built_lib, rows):
    """The fused FFN pair (ReLU backward folded into the down-projection's dgrad epilogue); gradients as small as the
    training loss produces them (1e-6): the power-of-two gradient scaling keeps them out of the fp16-subnormal range."""
    from fseend_b200.autograd import FfnFn, LinearFn
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, 256, generator=g).cuda().requires_grad_()
    w1 = (torch.randn(2048, 256, generator=g) / 16).cuda().requires_grad_()
    b1 = (0.1 * torch.randn(2048, generator=g)).cuda().requires_grad_()
    w2 = (torch.randn(256, 2048, generator=g) / 45).cuda().requires_grad_()
    b2 = (0.1 * torch.randn(256, generator=g)).cuda().requires_grad_()
    dy = (1e-6 * torch.randn(rows, 256, generator=g)).cuda()
    y = FfnFn.apply(x, w1, b1, w2, b2)
    y.backward(dy)
    ps = [t.detach().double().requires_grad_() for t in (x, w1, b1, w2, b2)]
    h = ps[0] @ ps[1].t() + ps[2]
    # gate with the native forward's sign pattern (pre-activations within rounding of zero would flip whole elements):
    # the same kernel on the same inputs reproduces the hidden layer bit for bit
    gate = (LinearFn.apply(x.detach(), w1.detach(), b1.detach(), "relu") > 0).double()
    yr = (h * gate) @ ps[3].t() + ps[4]
    yr.backward(dy.double())
    close(y, yr, "y")
    for t, r, name in zip((x, w1, b1, w2, b2), ps, ("dx", "dw1", "db1", "dw2", "db2")):
        close(t.grad, r.grad, name, rel=1e-4 if name in ("dx", "dw1", "db1") else REL)


def test_linear_no_input_grad_no_bias(built_lib):
    from fseend_b200.autograd import LinearFn
    x = torch.randn(130, 256).cuda()
    w = (torch.randn(128, 256) / 16).cuda().requires_grad_()
    y = LinearFn.apply(x, w, None, "none")
    y.sum().backward()
    close(y, x.double() @ w.detach().double().t(), "y")
    close(w.grad, torch.ones(130, 128).double().t() @ x.double().cpu(), "dw")


@pytest.mark.parametrize("rows,with_r", [(1000, True), (37, False)])
def test_add_layernorm_fwd_bwd(built_lib, rows, with_r):
    from fseend_b200.autograd import AddLayerNormFn
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, 256, generator=g).cuda().requires_grad_()
    r = (2 * torch.randn(rows, 256, generator=g)).cuda().requires_grad_() if with_r else None
    ga = (1 + 0.1 * torch.randn(256, generator=g)).cuda().requires_grad_()
    be = (0.1 * torch.randn(256, generator=g)).cuda().requires_grad_()
    dy = torch.randn(rows, 256, generator=g).cuda()
    y = AddLayerNormFn.apply(x, r, ga, be, 1e-5)
    y.backward(dy)
    xr, gr, br = (t.detach().double().requires_grad_() for t in (x, ga, be))
    rr = r.detach().double().requires_grad_() if with_r else None
    yr = nn.functional.layer_norm(xr + rr if with_r else xr, (256,), gr, br, 1e-5)
    yr.backward(dy.double())
    close(y, yr, "y")
    close(x.grad, xr.grad, "dx")
    if with_r:
        close(r.grad, rr.grad, "dr")
    close(ga.grad, gr.grad, "dgamma")
    close(be.grad, br.grad, "dbeta")


def ref_attention(qkv, delay):
    n, T, _ = qkv.shape
    q, k, v = (t.reshape(n, T, 4, 64).transpose(1, 2) for t in qkv.split(256, dim=-1))
    s = (q * 0.125) @ k.transpose(-1, -2)
    i = torch.arange(T, device=qkv.device)
    s = s.masked_fill(i[None, :] > i[:, None] + delay, float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n, T, 256)


@pytest.mark.parametrize("kernels", ["tensor", "cuda_core"])
@pytest.mark.parametrize("n,T,delay", [(3, 500, 0), (2, 130, 3), (1, 64, 0), (2, 1, 0), (1, 257, 70)])
def test_causal_attention_fwd_bwd(built_lib, monkeypatch, n, T, delay, kernels):
    """Both kernel families: mma.sync split-precision (default) and the fp32 CUDA-core form (FSEEND_TRAIN_ATTN=0)."""
    from fseend_b200.autograd import CausalAttnFn
    monkeypatch.setenv("FSEEND_TRAIN_ATTN", "1" if kernels == "tensor" else "0")
    g = torch.Generator().manual_seed(n * 1000 + T)
    qkv = (1.5 * torch.randn(n, T, 768, generator=g)).cuda().requires_grad_()
    do = torch.randn(n, T, 256, generator=g).cuda()
    o = CausalAttnFn.apply(qkv, delay)
    o.backward(do)
    qr = qkv.detach().double().requires_grad_()
    orf = ref_attention(qr, delay)
    orf.backward(do.double())
    close(o, orf, "out")
    close(qkv.grad, qr.grad, "dqkv")


@pytest.mark.parametrize("n,T,delay", [(4, 500, 0), (2, 77, 2)])
def test_encoder_layer_matches_torch(built_lib, n, T, delay):
    """The composed layer against the class the reference uses, gradients of every parameter and of the input."""
    from fseend_b200.autograd import encoder_layer_forward
    torch.manual_seed(5)
    layer = nn.TransformerEncoderLayer(256, 4, dim_feedforward=2048, dropout=0.0, batch_first=True).cuda()
    ref = nn.TransformerEncoderLayer(256, 4, dim_feedforward=2048, dropout=0.0, batch_first=True).double().cuda()
    ref.load_state_dict({k: v.double() for k, v in layer.state_dict().items()})
    x = torch.randn(n, T, 256).cuda().requires_grad_()
    dy = torch.randn(n, T, 256).cuda()
    y = encoder_layer_forward(layer, x, delay)
    y.backward(dy)
    i = torch.arange(T, device="cuda")
    mask = torch.zeros(T, T, device="cuda", dtype=torch.float64).masked_fill(i[None, :] > i[:, None] + delay, float("-inf"))
    xr = x.detach().double().requires_grad_()
    ref.train()           # the fused inference fast path is bypassed in train mode; dropout is 0
    yr = ref(xr, src_mask=mask)
    yr.backward(dy.double())
    close(y, yr, "y", rel=5e-5)
    close(x.grad, xr.grad, "dx", rel=5e-5)
    for (name, p), (_, pr) in zip(layer.named_parameters(), ref.named_parameters()):
        close(p.grad, pr.grad, name, rel=5e-5)


def test_rejects_cpu_tensors(built_lib):
    from fseend_b200.autograd import LinearFn
    from fseend_b200.native import FseendError
    with pytest.raises(FseendError):
        LinearFn.apply(torch.randn(4, 256), torch.randn(128, 256), None, "none")


@pytest.mark.parametrize("n,S", [(1000, 6), (37, 10), (5, 1), (64, 16)])
def test_speaker_attention_fwd_bwd(built_lib, n, S):
    from fseend_b200.autograd import SpeakerAttnFn
    g = torch.Generator().manual_seed(n + S)
    qkv = (1.5 * torch.randn(n, S, 768, generator=g)).cuda().requires_grad_()
    do = torch.randn(n, S, 256, generator=g).cuda()
    o = SpeakerAttnFn.apply(qkv)
    o.backward(do)
    qr = qkv.detach().double().requires_grad_()
    orf = ref_attention(qr, S)            # delay >= S: every key visible
    orf.backward(do.double())
    close(o, orf, "out")
    close(qkv.grad, qr.grad, "dqkv")


def _train_model(sd, enc_layers, dec_layers, mask_delay=0):
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=enc_layers,
                                       dec_n_layers=dec_layers, dropout=0.0, has_mask=True, max_seqlen=500,
                                       dec_dim_feedforward=2048, mask_delay=mask_delay)
    m.load_state_dict(sd, strict=True)
    return m.cuda().train()


@pytest.mark.parametrize("lens,n_spks,mask_delay", [([150, 97], [3, 2], 0), ([70], [4], 2)])
def test_training_forward_backward_matches_oracle_autograd(built_lib, lens, n_spks, mask_delay):
    """model(src, tgt, ilens) in train mode + standard_loss + emb loss (reference train/oln_tfm_enc_dec.py:78-85): values
    and the gradient of EVERY parameter against the float64 autograd of the oracle restatement.  BatchNorm is held in
    eval mode (running statistics) so that both sides normalise identically; batch-statistics mode is torch's own op."""
    from fseend_b200.loss import standard_loss
    from oracle import fs_eend_oracle as O
    sd = O.random_state_dict(seed=11, enc_n_layers=2, dec_n_layers=1)
    m = _train_model(sd, 2, 1, mask_delay)
    m.enc.bn.eval()
    src, _ = O.synthetic_features(len(lens), max(lens), seed=3, lens=lens)
    g = torch.Generator().manual_seed(9)
    tgt = [(torch.rand(l, n, generator=g) < 0.4).float() for l, n in zip(lens, n_spks)]
    out, emb_loss, embs, atts = m([s.cuda() for s in src], [t.cuda() for t in tgt], lens)
    assert [tuple(o.shape) for o in out] == [(l, n) for l, n in zip(lens, n_spks)]
    loss = standard_loss(out, [t.cuda() for t in tgt], label_delay=1) + emb_loss
    loss.backward()

    sd64 = {k: v.double().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe"))
            for k, v in sd.items()}
    cfg = O.Cfg(enc_n_layers=2, dec_n_layers=1, mask_delay=mask_delay)
    out_r, emb_loss_r, _, _ = O.forward(sd64, [s.double() for s in src], [t.double() for t in tgt], lens, cfg)
    bce = sum(torch.nn.functional.binary_cross_entropy_with_logits(y[1:], t.double()[:len(t) - 1]) * (len(y) - 1)
              for y, t in zip(out_r, tgt)) / (sum(lens) - len(lens))
    loss_r = bce + emb_loss_r
    loss_r.backward()
    assert abs(loss.item() - loss_r.item()) < 1e-5 * max(1.0, abs(loss_r.item()))
    for o, r in zip(out, out_r):
        close(o, r, "logits", rel=1e-4)
    # Yardstick: the same graph in plain float32 torch autograd (the arithmetic class of the reference).  Gradients of the
    # early layers are ill-conditioned (LayerNorm / ReLU chains amplify rounding: float32 torch itself is ~1e-3 off the
    # float64 value on some tensors), so each tensor must be within 3x the float32 error, or 2e-5 relative if that is larger.
    sd32 = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe"))
            for k, v in sd.items()}
    out_f, emb_loss_f, _, _ = O.forward(sd32, src, tgt, lens, cfg)
    (sum(torch.nn.functional.binary_cross_entropy_with_logits(y[1:], t[:len(t) - 1]) * (len(y) - 1)
         for y, t in zip(out_f, tgt)) / (sum(lens) - len(lens)) + emb_loss_f).backward()
    dead = {"dec.encoder.weight", "dec.encoder.bias", "dec.encoder_norm.weight", "dec.encoder_norm.bias"}
    checked, report = 0, []
    for name, p in m.named_parameters():
        r = sd64[name].grad
        if name in dead or "norm12" in name:
            assert p.grad is None and r is None, name        # parameters the forward never reads
            continue
        assert p.grad is not None and r is not None, name
        scale = r.abs().max().item()
        err = (p.grad.double().cpu() - r).abs().max().item() / scale
        err32 = (sd32[name].grad.double() - r).abs().max().item() / scale
        report.append((name, err, err32))
        assert err <= max(3 * err32, 2e-5), f"{name}: native {err:.2e} vs float32 torch {err32:.2e} (relative to max|grad|)"
        checked += 1
    assert checked >= 40


def test_training_step_reduces_loss(built_lib):
    """A few Adam steps through the drop-in model (batch-statistics BatchNorm, native forward/backward kernels) reduce the
    training loss on a fixed batch, and the inference path picks up the updated weights."""
    from fseend_b200.loss import standard_loss
    from oracle import fs_eend_oracle as O
    sd = O.random_state_dict(seed=4, enc_n_layers=2, dec_n_layers=1, trained_like=False)
    m = _train_model(sd, 2, 1)
    lens = [200, 160, 120]
    src = [s.cuda() for s in O.synthetic_features(3, 200, seed=8, lens=lens)[0]]
    g = torch.Generator().manual_seed(2)
    tgt = [(torch.rand(l, 3, generator=g) < 0.3).float().cuda() for l in lens]
    m.eval()
    y0, _, _ = m.test(src, lens, max_nspks=3)
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=5e-5)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        out, emb_loss, _, _ = m(src, tgt, lens)
        loss = standard_loss(out, tgt) + emb_loss
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses
    # trajectory of the same six steps in plain float32 torch on CPU (tests/test_train_graph_cpu.py stand-ins):
    # 0.8415 0.8071 0.7880 0.7794 0.7751 0.7716
    expect = [0.8415, 0.8071, 0.7880, 0.7794, 0.7751, 0.7716]
    assert max(abs(a - b) for a, b in zip(losses, expect)) < 2e-3, losses
    m.eval()
    y1, _, _ = m.test(src, lens, max_nspks=3)               # the inference pipeline picks up the updated weights
    assert all(torch.isfinite(o).all() for o in y1)
    assert max((a - b).abs().max().item() for a, b in zip(y0, y1)) > 1e-3


def keep_mask(seed, shape, p):
    """The kernels' dropout hash (csrc/train_attn.cu drop_hash) in torch integer arithmetic: True = kept."""
    import numpy as np
    n = 1
    for d in shape:
        n *= d
    M = 0xFFFFFFFF
    idx = torch.arange(n, dtype=torch.int64)
    x = (idx & M) ^ (seed & M)
    x = (x * 0x9E3779B1) & M
    x = x ^ (x >> 15)
    x = (x + (idx >> 32) * 0x85EBCA77 + (seed >> 32)) & M
    x = (x * 0xC2B2AE3D) & M
    x = x ^ (x >> 13)
    x = (x * 0x27D4EB2F) & M
    x = x ^ (x >> 16)
    thresh = min(2 ** 32 - 1, int(float(np.float32(p)) * 4294967296.0))
    return (x >= thresh).reshape(shape)


def ref_attention_dropout(qkv, delay, keep, p):
    n, T, _ = qkv.shape
    q, k, v = (t.reshape(n, T, 4, 64).transpose(1, 2) for t in qkv.split(256, dim=-1))
    s = (q * 0.125) @ k.transpose(-1, -2)
    i = torch.arange(T, device=qkv.device)
    s = s.masked_fill(i[None, :] > i[:, None] + delay, float("-inf"))
    pr = torch.softmax(s, -1) * keep.to(qkv) / (1 - float(torch.tensor(p, dtype=torch.float32)))
    return (pr @ v).transpose(1, 2).reshape(n, T, 256)


@pytest.mark.parametrize("kernels", ["tensor", "cuda_core"])
def test_causal_attention_interleaved_layout(built_lib, monkeypatch, kernels):
    """The attractor decoder's [B, T, S, 768] tensor read in place (sequence (b, s) with row stride S), with dropout: equal
    to the plain layout on the transposed copy, forward and backward, same mask (sequence index b * S + s)."""
    from fseend_b200.autograd import CausalAttnFn
    monkeypatch.setenv("FSEEND_TRAIN_ATTN", "1" if kernels == "tensor" else "0")
    g = torch.Generator().manual_seed(31)
    B, T, S = 2, 150, 3
    q4 = (1.5 * torch.randn(B, T, S, 768, generator=g)).cuda().requires_grad_()
    do4 = torch.randn(B, T, S, 256, generator=g).cuda()
    o4 = CausalAttnFn.apply(q4, 1, 0.2, 4242)
    o4.backward(do4)
    q3 = q4.detach().transpose(1, 2).reshape(B * S, T, 768).contiguous().requires_grad_()
    o3 = CausalAttnFn.apply(q3, 1, 0.2, 4242)
    o3.backward(do4.transpose(1, 2).reshape(B * S, T, 256).contiguous())
    assert torch.equal(o4.detach().transpose(1, 2).reshape(B * S, T, 256), o3.detach())
    assert torch.equal(q4.grad.transpose(1, 2).reshape(B * S, T, 768), q3.grad)


def test_causal_attention_tiny_gradients(built_lib):
    """Upstream gradients as the training loss produces them (1e-7, far below the fp16 normal range): the tensor-core
    backward rescales dO by a power of two before the operand split."""
    from fseend_b200.autograd import CausalAttnFn
    g = torch.Generator().manual_seed(77)
    qkv = (1.5 * torch.randn(2, 300, 768, generator=g)).cuda().requires_grad_()
    do = (1e-7 * torch.randn(2, 300, 256, generator=g) * torch.rand(2, 300, 1, generator=g) ** 4).cuda()
    CausalAttnFn.apply(qkv, 0).backward(do)
    qr = qkv.detach().double().requires_grad_()
    ref_attention(qr, 0).backward(do.double())
    close(qkv.grad, qr.grad, "dqkv")


@pytest.mark.parametrize("kernels", ["tensor", "cuda_core"])
@pytest.mark.parametrize("n,T,delay,p", [(2, 200, 0, 0.1), (1, 70, 3, 0.5)])
def test_causal_attention_dropout(built_lib, monkeypatch, n, T, delay, p, kernels):
    """Attention-probability dropout: forward and backward against a torch reference that applies the SAME mask
    (regenerated from the seed by the hash above); the keep rate matches 1 - p."""
    from fseend_b200.autograd import CausalAttnFn
    monkeypatch.setenv("FSEEND_TRAIN_ATTN", "1" if kernels == "tensor" else "0")
    seed = 0x1234567 + (T << 33)
    g = torch.Generator().manual_seed(T)
    qkv = (1.5 * torch.randn(n, T, 768, generator=g)).cuda().requires_grad_()
    do = torch.randn(n, T, 256, generator=g).cuda()
    o = CausalAttnFn.apply(qkv, delay, p, seed)
    o.backward(do)
    keep = keep_mask(seed, (n, 4, T, T), p).cuda()
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    qr = qkv.detach().double().requires_grad_()
    orf = ref_attention_dropout(qr, delay, keep, p)
    orf.backward(do.double())
    close(o, orf, "out")
    close(qkv.grad, qr.grad, "dqkv")
    o0 = CausalAttnFn.apply(qkv.detach(), delay, 0.0, seed)
    assert (o0 - o.detach()).abs().max() > 1e-3            # dropout did change the output


@pytest.mark.parametrize("n,S,p", [(300, 6, 0.1), (33, 10, 0.4)])
def test_speaker_attention_dropout(built_lib, n, S, p):
    from fseend_b200.autograd import SpeakerAttnFn
    seed = 987654321012345
    g = torch.Generator().manual_seed(n)
    qkv = (1.5 * torch.randn(n, S, 768, generator=g)).cuda().requires_grad_()
    do = torch.randn(n, S, 256, generator=g).cuda()
    o = SpeakerAttnFn.apply(qkv, p, seed)
    o.backward(do)
    keep = keep_mask(seed, (n, 4, S, S), p).cuda()
    qr = qkv.detach().double().requires_grad_()
    orf = ref_attention_dropout(qr, S, keep, p)
    orf.backward(do.double())
    close(o, orf, "out")
    close(qkv.grad, qr.grad, "dqkv")


def test_training_with_reference_dropout(built_lib):
    """The reference recipes train with dropout 0.1: the step runs, is reproducible under torch.manual_seed, differs
    between seeds, and eval mode is unaffected."""
    from fseend_b200.loss import standard_loss
    from oracle import fs_eend_oracle as O
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    sd = O.random_state_dict(seed=4, enc_n_layers=1, dec_n_layers=1, trained_like=False)
    m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=1, dec_n_layers=1,
                                       dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048)
    m.load_state_dict(sd)
    m = m.cuda().train()
    lens = [120, 90]
    src = [s.cuda() for s in O.synthetic_features(2, 120, seed=8, lens=lens)[0]]
    tgt = [(torch.rand(l, 3) < 0.3).float().cuda() for l in lens]

    def run(seed):
        torch.manual_seed(seed)
        m.zero_grad()
        out, emb_loss, _, _ = m(src, tgt, lens)
        loss = standard_loss(out, tgt) + emb_loss
        loss.backward()
        return loss.item(), m.enc.encoder.weight.grad.clone()

    l1, g1 = run(1)
    l2, g2 = run(1)
    l3, g3 = run(2)
    assert l1 == l2 and torch.equal(g1, g2)
    assert l1 != l3 and torch.isfinite(g3).all()


def test_l2norm_head_batchnorm_fwd_bwd(built_lib):
    from fseend_b200.autograd import BatchNormTrainFn, HeadFn, L2NormFn
    g = torch.Generator().manual_seed(21)
    # L2 normalisation
    x = (3 * torch.randn(5, 37, 256, generator=g)).cuda().requires_grad_()
    dy = torch.randn(5, 37, 256, generator=g).cuda()
    y = L2NormFn.apply(x)
    y.backward(dy)
    xr = x.detach().double().requires_grad_()
    yr = xr / torch.norm(xr, dim=-1, keepdim=True)
    yr.backward(dy.double())
    close(y, yr, "l2 y")
    close(x.grad, xr.grad, "l2 dx")
    # head
    emb = torch.randn(3, 50, 256, generator=g).cuda().requires_grad_()
    att = torch.randn(3, 50, 6, 256, generator=g).cuda().requires_grad_()
    dl = torch.randn(3, 50, 6, generator=g).cuda()
    lg = HeadFn.apply(emb, att)
    lg.backward(dl)
    er, ar = emb.detach().double().requires_grad_(), att.detach().double().requires_grad_()
    lr = (er[:, :, None, :] * ar).sum(-1)
    lr.backward(dl.double())
    close(lg, lr, "head y")
    close(emb.grad, er.grad, "head demb")
    close(att.grad, ar.grad, "head datt")
    # BatchNorm, training mode, C = 345 (not a multiple of 4), rows not a multiple of the block
    xb = (2 * torch.randn(7, 300, 345, generator=g) + 0.5).cuda().requires_grad_()
    gam = (1 + 0.2 * torch.randn(345, generator=g)).cuda().requires_grad_()
    bet = (0.1 * torch.randn(345, generator=g)).cuda().requires_grad_()
    dyb = torch.randn(7, 300, 345, generator=g).cuda()
    yb, stats = BatchNormTrainFn.apply(xb, gam, bet, 1e-5)
    yb.backward(dyb)
    xr, gr, br = (t.detach().double().requires_grad_() for t in (xb, gam, bet))
    ybr = torch.nn.functional.batch_norm(xr.reshape(-1, 345), None, None, gr, br, True, 0.1, 1e-5).reshape(7, 300, 345)
    ybr.backward(dyb.double())
    close(yb, ybr, "bn y")
    close(stats[:345], xr.detach().reshape(-1, 345).mean(0), "bn mean")
    close(stats[345:], xr.detach().reshape(-1, 345).var(0, unbiased=False), "bn var")
    close(xb.grad, xr.grad, "bn dx", rel=1e-4)
    close(gam.grad, gr.grad, "bn dgamma")
    close(bet.grad, br.grad, "bn dbeta")


def test_training_batchnorm_train_mode_matches_torch_graph(built_lib, monkeypatch):
    """The whole step with BatchNorm in training mode (batch statistics over the -1-padded batch, running statistics
    updated) against the same graph built from plain torch ops in float64 on the GPU (the stand-ins of
    tests/test_train_graph_cpu.py, which that file pins against the oracle)."""
    import copy
    import fseend_b200.autograd as A
    import fseend_b200.train_graph as G
    import test_train_graph_cpu as S
    from fseend_b200.loss import standard_loss
    from oracle import fs_eend_oracle as O
    sd = O.random_state_dict(seed=12, enc_n_layers=1, dec_n_layers=1)
    m = _train_model(sd, 1, 1)
    ref = copy.deepcopy(m).double()
    lens, n_spks = [130, 88, 61], [3, 2, 3]
    src, _ = O.synthetic_features(3, 130, seed=5, lens=lens)
    g = torch.Generator().manual_seed(6)
    tgt = [(torch.rand(l, n, generator=g) < 0.4).float().cuda() for l, n in zip(lens, n_spks)]
    out, el, _, _ = m([s.cuda() for s in src], tgt, lens)
    (standard_loss(out, tgt) + el).backward()
    with monkeypatch.context() as mp:
        for mod in (A, G):
            mp.setattr(mod, "LinearFn", S._Lin)
            mp.setattr(mod, "AddLayerNormFn", S._AddLn)
        mp.setattr(A, "FfnFn", S._Ffn)
        mp.setattr(A, "CausalAttnFn", S._Causal)
        mp.setattr(A, "SpeakerAttnFn", type("P", (), {"apply": staticmethod(lambda qkv, p=0.0, seed=0: ref_attention(qkv, 1 << 20))}))
        mp.setattr(G, "L2NormFn", S._L2)
        mp.setattr(G, "HeadFn", S._Head)
        mp.setattr(G, "batch_norm_forward", S._bn)
        out_r, el_r, _, _ = G.fs_forward_train(ref, [s.cuda().double() for s in src], [t.double() for t in tgt], lens)
        (G.standard_loss_train(out_r, [t.double() for t in tgt]) + el_r).backward()
    for o, r in zip(out, out_r):
        close(o, r, "logits", rel=1e-4)
    close(m.enc.bn.running_mean, ref.enc.bn.running_mean, "running_mean")
    close(m.enc.bn.running_var, ref.enc.bn.running_var, "running_var")
    assert int(m.enc.bn.num_batches_tracked) == int(ref.enc.bn.num_batches_tracked) == 1
    for (name, p), (_, pr) in zip(m.named_parameters(), ref.named_parameters()):
        if pr.grad is None:
            assert p.grad is None, name
            continue
        close(p.grad, pr.grad, name, rel=3e-4)


@pytest.mark.parametrize("case", ["train_e2d1", "train_e1d2_delay"])
def test_training_step_matches_real_reference_gradients(built_lib, case):
    """model(src, tgt, ilens) -> standard_loss + emb loss -> backward through the native kernels (BatchNorm with batch
    statistics, label delay, mask delay) against gradients produced by the REAL reference's autograd
    (tests/golden/make_golden_train.py): losses, logits checksum, running statistics, and for every parameter the gradient
    norm and 24 sampled elements, within 1e-3 (relative to the norm / to max|grad|)."""
    import test_train_graph_cpu as S
    from fseend_b200.loss import standard_loss
    from oracle import fs_eend_oracle as O
    wseed, ne, nd, lens, n_spks, md, ld = S.TRAIN_CASES[case]
    sd = O.random_state_dict(seed=wseed, enc_n_layers=ne, dec_n_layers=nd)
    m = _train_model(sd, ne, nd, md)
    src, _ = O.synthetic_features(len(lens), max(lens), seed=wseed, lens=lens)
    tgt = [t.cuda() for t in O.synthetic_labels(wseed, lens, n_spks)]
    out, el, _, _ = m([s.cuda() for s in src], tgt, lens)
    bce = standard_loss(out, tgt, label_delay=ld)
    (bce + el).backward()
    S.check_against_reference_grads(m, S.load_train_golden()[case], bce, el, out, rel=1e-3)
