"""Per-kernel parity on a B200: each sm_100a kernel (called through the C ABI) against plain torch fp32
arithmetic on the SAME fp16-rounded operands.  Tolerances: outputs are stored as fp16 (rel. 2^-11), fp32
accumulation order differs -> max-abs 5e-3 on O(1) values; logic errors are O(1)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def ln_ref(x, g, b, eps=1e-5):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)


@pytest.fixture(scope="module")
def N():
    from fseend_b200 import native
    native.lib()
    assert torch.cuda.is_available() and native.lib().fseend_device_ok() == 1
    return native


def test_gemm_bias_relu_ragged_rows(N):
    a = rnd(300, 256, seed=1).half()
    w = rnd(768, 256, scale=1 / 16, seed=2).half()
    bias = rnd(768, seed=3)
    out = N.op_gemm(a, w, N.EPI_BIAS, bias=bias)
    ref = a.float() @ w.float().T + bias
    assert (out.float() - ref).abs().max().item() < 5e-3
    out = N.op_gemm(a, w, N.EPI_BIAS, bias=bias, relu=True)
    assert (out.float() - ref.relu()).abs().max().item() < 5e-3


def test_gemm_k384_layernorm(N):
    a = rnd(1000, 384, seed=4).half()
    w = rnd(256, 384, scale=1 / 20, seed=5).half()
    bias, g, b = rnd(256, seed=6), 1 + 0.3 * rnd(256, seed=7), 0.1 * rnd(256, seed=8)
    out = N.op_gemm(a, w, N.EPI_LN, bias=bias, ln_g=g, ln_b=b)
    ref = ln_ref(a.float() @ w.float().T + bias, g, b)
    assert (out.float() - ref).abs().max().item() < 5e-3


def test_gemm_k2048_residual_layernorm_large_mean(N):
    a = rnd(517, 2048, seed=9).half()
    w = rnd(256, 2048, scale=1 / 45, seed=10).half()
    res = (rnd(517, 256, seed=11) + 3.0).half()          # large mean: exercises the robust variance merge
    bias, g, b = rnd(256, seed=12), 1 + 0.3 * rnd(256, seed=13), 0.1 * rnd(256, seed=14)
    out = N.op_gemm(a, w, N.EPI_LN, bias=bias, residual=res, ln_g=g, ln_b=b)
    ref = ln_ref(a.float() @ w.float().T + bias + res.float(), g, b)
    assert (out.float() - ref).abs().max().item() < 5e-3


def test_gemm_layernorm_zero_rows_beyond_len(N):
    n_seq, T = 3, 150
    a = rnd(n_seq * T, 256, seed=15).half()
    w = rnd(256, 256, scale=1 / 16, seed=16).half()
    g, b = 1 + 0.3 * rnd(256, seed=17), 0.1 * rnd(256, seed=18)
    lens = torch.tensor([150, 77, 1], dtype=torch.int32, device=DEV)
    out = N.op_gemm(a, w, N.EPI_LN, n_seq=n_seq, ln_g=g, ln_b=b, seq_len=lens).view(n_seq, T, 256)
    ref = ln_ref(a.float() @ w.float().T, g, b).view(n_seq, T, 256)
    for i, l in enumerate(lens.tolist()):
        assert (out[i, :l].float() - ref[i, :l]).abs().max().item() < 5e-3
        assert out[i, l:].abs().max().item() == 0 if l < T else True


def test_gemm_conv_taps_l2(N):
    """Conv1d(256,256,19,padding=9) + L2 norm as 19 shifted GEMMs with per-sequence zero fill."""
    n_seq, T, K = 3, 200, 19
    x = rnd(n_seq, T, 256, seed=19).half()
    wc = rnd(256, 256, K, scale=1 / 70, seed=20)                    # (out, in, k)
    bias = rnd(256, seed=21) * 0.1
    w_taps = wc.permute(2, 0, 1).contiguous().half()                 # [k][out][in]
    out = N.op_gemm(x.view(-1, 256), w_taps.view(K * 256, 256), N.EPI_L2, n_seq=n_seq, taps=K, tap_shift=-9,
                    bias=bias).view(n_seq, T, 256)
    y = torch.nn.functional.conv1d(x.float().transpose(1, 2), w_taps.float().permute(1, 2, 0), bias, padding=9)
    y = y.transpose(1, 2)
    ref = y / y.norm(dim=-1, keepdim=True)
    assert (out.float() - ref).abs().max().item() < 2e-3


def test_gemm_convert_broadcast(N):
    S = 6
    a = rnd(333, 256, seed=22).half()
    w = rnd(256, 256, scale=1 / 16, seed=23).half()
    pe = rnd(16, 256, seed=24)
    out = N.op_gemm(a, w, N.EPI_CONVERT, pe_proj=pe, S=S)
    ref = (a.float() @ w.float().T)[:, None, :] + pe[None, :S, :]
    assert out.shape == (333, S, 256)
    assert (out.float() - ref).abs().max().item() < 5e-3


def test_gemm_k256_residual_layernorm_odd_tiles(N):
    a = rnd(128 * 7 + 5, 256, seed=25).half()                # 8 row tiles -> 4 pairs; 7*128+5 rows: ragged last tile
    w = rnd(256, 256, scale=1 / 16, seed=26).half()
    res = rnd(128 * 7 + 5, 256, seed=27).half()
    bias, g, b = rnd(256, seed=28), 1 + 0.3 * rnd(256, seed=29), 0.1 * rnd(256, seed=30)
    out = N.op_gemm(a, w, N.EPI_LN, bias=bias, residual=res, ln_g=g, ln_b=b)
    ref = ln_ref(a.float() @ w.float().T + bias + res.float(), g, b)
    assert (out.float() - ref).abs().max().item() < 5e-3
    a3 = a[: 128 * 3 - 1]                                     # odd tile count: the pair kernel's last CTA idles
    out = N.op_gemm(a3, w, N.EPI_LN, bias=bias, residual=res[: a3.shape[0]], ln_g=g, ln_b=b)
    assert (out.float() - ref[: a3.shape[0]]).abs().max().item() < 5e-3


def test_gemm_many_tiles_wide(N):
    """More row-tile pairs than clusters and three n-tiles: every cluster loops over several items."""
    a = rnd(128 * 201, 256, seed=31).half()
    w = rnd(768, 256, scale=1 / 16, seed=32).half()
    bias = rnd(768, seed=33)
    out = N.op_gemm(a, w, N.EPI_BIAS, bias=bias)
    ref = a.float() @ w.float().T + bias
    assert (out.float() - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("case", ["bias_relu_ragged_rows", "layernorm_zero_rows_beyond_len", "convert_broadcast",
                                  "k256_residual_layernorm_odd_tiles", "many_tiles_wide", "k384_layernorm"])
def test_gemm_pair_kernel_forced(N, monkeypatch, case):
    """FSEEND_GEMM_PAIR=2 routes every structurally eligible GEMM (one tap, K <= 256) through the weight-stationary
    CTA-pair kernel, whatever its size; K = 384 checks the fallback."""
    monkeypatch.setenv("FSEEND_GEMM_PAIR", "2")
    globals()["test_gemm_" + case](N)


def attn_ref(qkv, mask_delay, scale=0.125):
    B, T, S, _ = qkv.shape
    x = qkv.float().permute(0, 2, 1, 3).reshape(B * S, T, 3, 4, 64)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * scale
    i = torch.arange(T, device=qkv.device)
    s = s.masked_fill(i[None, :] > i[:, None] + mask_delay, float("-inf"))
    o = torch.softmax(s, dim=-1) @ v                                   # (B*S, H, T, 64)
    return o.transpose(1, 2).reshape(B, S, T, 256).permute(0, 2, 1, 3)


ATTN_SHAPES = [(2, 300, 1, 0), (1, 500, 3, 0), (2, 130, 2, 2), (1, 64, 1, 0), (1, 257, 1, 300), (1, 700, 2, 0),
               (3, 1000, 1, 5), (1, 129, 1, 0)]


@pytest.mark.parametrize("step", ["128", "256"])
@pytest.mark.parametrize("B,T,S,md", ATTN_SHAPES)
def test_causal_attention(N, monkeypatch, B, T, S, md, step):
    """attn3.cu (the decoder's kernel: one self-contained warpgroup per query tile, whole-step softmax, dynamic item
    queue; 128- and 256-key steps; odd and even tile counts, partial last tiles, look-ahead, no-mask, multi-step rows
    at T = 700 / 1000)."""
    monkeypatch.setenv("FSEEND_ATTN", "3")
    monkeypatch.setenv("FSEEND_ATTN_STEP", step)
    qkv = rnd(B, T, S, 768, seed=25 + T).half()
    out = N.op_causal_attn(qkv, mask_delay=md)
    ref = attn_ref(qkv, md)
    assert (out.float() - ref).abs().max().item() < 4e-3


@pytest.mark.parametrize("B,T,S,md", ATTN_SHAPES)
def test_causal_attention_tile_pair_kernel(N, monkeypatch, B, T, S, md):
    """FSEEND_ATTN=2: the round-1 default (attn2.cu: pairs of query tiles, role warps)."""
    monkeypatch.setenv("FSEEND_ATTN", "2")
    qkv = rnd(B, T, S, 768, seed=25 + T).half()
    out = N.op_causal_attn(qkv, mask_delay=md)
    assert (out.float() - attn_ref(qkv, md)).abs().max().item() < 4e-3


@pytest.mark.parametrize("B,T,S,md", ATTN_SHAPES)
def test_causal_attention_one_tile_kernel(N, monkeypatch, B, T, S, md):
    """FSEEND_ATTN=1: the one-query-tile-per-item kernel (attn.cu)."""
    monkeypatch.setenv("FSEEND_ATTN", "1")
    qkv = rnd(B, T, S, 768, seed=25 + T).half()
    out = N.op_causal_attn(qkv, mask_delay=md)
    assert (out.float() - attn_ref(qkv, md)).abs().max().item() < 4e-3


@pytest.mark.parametrize("variant", ["", "1", "2", "3"])
def test_causal_attention_large_scores_exercise_lazy_rescale(N, monkeypatch, variant):
    """Scores with a spread of ~±60 in log2 units: the running maximum outgrows the lazy reference by more than 2^8
    many times per row, so the rare path (rescale O in TMEM, the row sum and the P chunks already written) runs."""
    if variant:
        monkeypatch.setenv("FSEEND_ATTN", variant)
    qkv = rnd(2, 400, 2, 768, seed=77)
    qkv[..., :512] *= 4.0                      # q and k: scores ~ N(0, 16^2) before the 1/8 scale
    qkv = qkv.half()
    out = N.op_causal_attn(qkv)
    ref = attn_ref(qkv, 0)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("S", [4, 6, 10, 16])
def test_speaker_attention(N, S):
    F = 777
    qkv = rnd(F, S, 768, seed=40 + S).half()
    out = N.op_spk_attn(qkv)
    x = qkv.float().view(F, S, 3, 4, 64)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    o = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
    ref = o.transpose(1, 2).reshape(F, S, 256)
    assert (out.float() - ref).abs().max().item() < 3e-3


def test_head(N):
    F, S = 1001, 6
    emb = torch.nn.functional.normalize(rnd(F, 256, seed=50), dim=-1).half()
    att = rnd(F, S, 256, seed=51).half()
    logits, e32, a32 = N.op_head(emb, att, want_f32=True)
    an = att.float() / att.float().norm(dim=-1, keepdim=True)
    ref = (emb.float()[:, None, :] * an).sum(-1)
    assert (logits - ref).abs().max().item() < 1e-5
    assert (a32 - an).abs().max().item() < 1e-6
    assert torch.equal(e32, emb.float())


def test_prep_input(N):
    lens = [50, 17, 33]
    B, T, Din, Kpad = 3, 50, 345, 384
    x = rnd(sum(lens), Din, seed=52)
    cu = torch.tensor([0, 50, 67, 100], dtype=torch.int32, device=DEV)
    sc, sh = 1 + 0.2 * rnd(Din, seed=53), rnd(Din, seed=54)
    out = N.op_prep_input(x, cu, B, T, Kpad, sc, sh)
    ref = torch.zeros(B, T, Kpad, device=DEV)
    off = 0
    for b, l in enumerate(lens):
        ref[b, :l, :Din] = x[off:off + l] * sc + sh
        ref[b, l:, :Din] = -1.0 * sc + sh
        off += l
    assert torch.equal(out, ref.half())


@pytest.mark.parametrize("S", [4, 6, 10, 16])
def test_speaker_attention_tensor_core(N, S):
    """Block-diagonal tcgen05 variant (the one the model uses) vs fp32 torch."""
    F = 777
    qkv = rnd(F, S, 768, seed=40 + S).half()
    out = N.op_spk_attn(qkv, tensor_core=True)
    x = qkv.float().view(F, S, 3, 4, 64)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    o = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
    ref = o.transpose(1, 2).reshape(F, S, 256)
    assert (out.float() - ref).abs().max().item() < 4e-3


def ffn_ref(x, w1, b1, w2, b2, g, b):
    h = torch.relu(x.float() @ w1.float().T + b1).half().float()      # hidden activations are fp16 on chip
    return ln_ref(x.float() + h @ w2.float().T + b2, g, b)


@pytest.mark.parametrize("cluster", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("rows,F", [(300, 2048), (128 * 5, 256), (1, 1024)])
def test_fused_ffn(N, cluster, rows, F):
    x = rnd(rows, 256, seed=60).half()
    w1 = rnd(F, 256, scale=1 / 16, seed=61).half()
    w2 = rnd(256, F, scale=1 / math.sqrt(F), seed=62).half()
    b1, b2 = rnd(F, seed=63) * 0.5, rnd(256, seed=64) * 0.5
    g, b = 1 + 0.3 * rnd(256, seed=65), 0.1 * rnd(256, seed=66)
    out = N.op_ffn(x, w1, b1, w2, b2, g, b, cluster=cluster)
    ref = ffn_ref(x, w1, b1, w2, b2, g, b)
    assert (out.float() - ref).abs().max().item() < 6e-3


@pytest.mark.parametrize("cluster", [4, 5])
def test_fused_ffn_per_sequence_zero_rows(N, cluster):
    n_seq, T, F = 3, 150, 512
    x = rnd(n_seq * T, 256, seed=70).half()
    w1 = rnd(F, 256, scale=1 / 16, seed=71).half()
    w2 = rnd(256, F, scale=1 / 22, seed=72).half()
    b1, b2 = rnd(F, seed=73) * 0.5, rnd(256, seed=74) * 0.5
    g, b = 1 + 0.3 * rnd(256, seed=75), 0.1 * rnd(256, seed=76)
    lens = torch.tensor([150, 77, 1], dtype=torch.int32, device=DEV)
    out = N.op_ffn(x, w1, b1, w2, b2, g, b, n_seq=n_seq, seq_len=lens, cluster=cluster).view(n_seq, T, 256)
    ref = ffn_ref(x, w1, b1, w2, b2, g, b).view(n_seq, T, 256)
    for i, l in enumerate(lens.tolist()):
        assert (out[i, :l].float() - ref[i, :l]).abs().max().item() < 6e-3
        if l < T:
            assert out[i, l:].abs().max().item() == 0


def _embloss_ref(emb, labels, lens=None):
    """torch fp64 restatement of FS:model:46-57 (lens None) / LS:model:92-113 (masked) on the given tensors."""
    e, l = emb.double(), labels.double()
    if lens is not None:
        mask = (torch.arange(e.shape[1], device=e.device)[None, :] < torch.as_tensor(lens, device=e.device)[:, None])
        e, l = e * mask[..., None], l * mask[..., None]
    en, ln = e.norm(dim=-1, keepdim=True), l.norm(dim=-1, keepdim=True)
    amap = e @ e.transpose(-1, -2) / (en @ en.transpose(-1, -2) + 1e-6)
    lmap = l @ l.transpose(-1, -2) / (ln @ ln.transpose(-1, -2) + 1e-6)
    sq = ((amap - lmap) ** 2).sum()
    return (sq / (sum(int(x) ** 2 for x in lens) if lens is not None else e.shape[0] * e.shape[1] ** 2)).item()


@pytest.mark.parametrize("B,T,S,masked", [(3, 500, 6, False), (2, 128, 4, False), (4, 333, 10, True), (1, 1, 1, False),
                                          (2, 700, 16, True), (5, 129, 5, True)])
def test_embloss_kernel(N, B, T, S, masked):
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
    emb = torch.nn.functional.normalize(torch.randn(B, T, 256, generator=g), dim=-1).to(DEV)
    labels = (torch.rand(B, T, S, generator=g) > 0.6).float().to(DEV)
    lens = None
    if masked:
        lens = [T] + [max(1, T - 37 * (b + 1)) for b in range(B - 1)]
        for b, l in enumerate(lens):
            labels[b, l:] = 0
    seq = torch.tensor(lens, device=DEV, dtype=torch.int32) if masked else None
    div = float(sum(l * l for l in lens)) if masked else None
    loss = N.op_embloss(emb, labels, seq_len=seq, divisor=div)
    loss2 = N.op_embloss(emb, labels, seq_len=seq, divisor=div)
    ref = _embloss_ref(emb, labels, lens)
    assert loss.item() == loss2.item()                      # fixed-order reduction: bit-reproducible
    assert abs(loss.item() - ref) < 2e-4 * max(1.0, abs(ref)), (loss.item(), ref)


def test_embloss_rejects_bad_arguments(N):
    emb = torch.zeros(1, 8, 256, device=DEV)
    with pytest.raises(N.FseendError):
        N.op_embloss(emb, torch.zeros(1, 8, 17, device=DEV))          # S > 16
    with pytest.raises(N.FseendError):
        N.op_embloss(emb.half(), torch.zeros(1, 8, 4, device=DEV))    # fp32 only


@pytest.mark.parametrize("frames,S", [(500, 6), (37, 4), (300, 10), (9, 16), (130, 5), (1, 1)])
def test_spk_qkv_attn_fused(N, frames, S):
    """Fused QKV projection + speaker-axis attention vs torch fp32 on the same fp16 operands (q, k, v stay fp32 inside
    the kernel, so it is compared against the unrounded projection)."""
    x = rnd(frames, S, 256, seed=frames).half()
    w = rnd(768, 256, scale=1 / 16, seed=2).half()
    b = rnd(768, seed=3) * 0.2
    out = N.op_spk_qkv_attn(x, w, b)
    qkv = x.float() @ w.float().T + b
    q, k, v = (t.reshape(frames, S, 4, 64).transpose(1, 2) for t in qkv.split(256, dim=-1))
    att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
    ref = att.transpose(1, 2).reshape(frames, S, 256)
    assert out.shape == ref.shape
    assert (out.float() - ref).abs().max().item() < 4e-3
