"""End-to-end parity of the sm_100a FS-EEND forward (through the nnet API mirror and the C ABI) against
(i) golden logits produced by the real reference and (ii) the CPU oracle, plus size-independent properties
at the BASELINE.json configuration (B=64, T=500, S=6).  Tolerance: 1e-3 max-abs on logits (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import fs_eend_oracle as O
from test_oracle import CASES, WIDE_CASES, load_case, load_wide_case

pytestmark = pytest.mark.gpu
TOL = 1e-3


def make_model(sd, mask_delay=0):
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4,
                                       dec_n_layers=2, dropout=0.1, has_mask=True, max_seqlen=500,
                                       dec_dim_feedforward=2048, mask_delay=mask_delay)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("name", list(CASES))
def test_logits_match_reference_golden(name):
    sd, src, lens, S, cfg, g = load_case(name)
    m = make_model(sd, cfg.mask_delay)
    out, emb, att = m.test([s.cuda() for s in src], lens, max_nspks=S)
    worst = 0.0
    for i, o in enumerate(out):
        ref = g[f"logits_{i}"]
        assert tuple(o.shape) == ref.shape
        worst = max(worst, float(np.abs(o.cpu().numpy() - ref).max()))
        stride = int(g["emb_stride"][i])
        assert np.abs(emb[i].cpu().numpy()[::stride] - g[f"emb_{i}"]).max() < TOL
        assert att[i].shape == (lens[i], S, 256)
    print(f"{name}: max-abs logit error vs reference = {worst:.2e}")
    assert worst < TOL


@pytest.mark.parametrize("name", list(WIDE_CASES))
def test_wide_dynamic_range_logits_and_sigmoid_space_check(name):
    """Weights whose logits span most of the cosine range (std 0.26 instead of 0.02): the 1e-3 absolute bound is then
    0.4 % of the signal.  Also the reference's own acceptance check, in sigmoid space with atol = rtol = 1e-4
    (FS-EEND/streaming_infer_dia.py:97), against the real reference's logits."""
    sd, src, lens, S, g = load_wide_case(name)
    m = make_model(sd)
    out, _, _ = m.test([s.cuda() for s in src], lens, max_nspks=S)
    worst = 0.0
    for i, o in enumerate(out):
        ref = torch.from_numpy(g[f"logits_{i}"])
        worst = max(worst, float((o.cpu() - ref).abs().max()))
        assert torch.allclose(torch.sigmoid(o.cpu()[:, 1:]), torch.sigmoid(ref[:, 1:]), atol=1e-4, rtol=1e-4)
    print(f"{name}: max-abs logit error vs reference = {worst:.2e} on logits in "
          f"[{min(float(g[f'logits_{i}'].min()) for i in range(len(lens))):.2f}, "
          f"{max(float(g[f'logits_{i}'].max()) for i in range(len(lens))):.2f}]")
    assert worst < TOL


def test_forward_api_and_emb_consistency_loss():
    sd, src, lens, S, cfg, g = load_case("ragged_S6")
    m = make_model(sd)
    gen = torch.Generator().manual_seed(123)
    tgt = [(torch.rand(l, S, generator=gen) > 0.6).float() for l in lens]
    out, loss, emb, att = m(([s.cuda() for s in src]), tgt, lens)
    assert abs(loss.item() - float(g["emb_consis_loss"])) < 1e-3
    assert np.abs(out[0].cpu().numpy() - g["fwd_logits_0"]).max() < TOL
    assert att[0].shape == (lens[0], S - 1, 256)


def test_host_buffer_entry_point_matches_device_entry_point():
    sd, src, lens, S, cfg, g = load_case("ragged_S6")
    m = make_model(sd)
    x = torch.cat(src).contiguous()
    dev_logits, _, _ = m.native().forward(x.cuda(), lens, S)
    host_logits, _, _ = m.native().forward_host(x.pin_memory(), lens, S)
    assert torch.equal(dev_logits.cpu(), host_logits)


def test_pipelined_host_entry_point_matches_blocking_call():
    """fseend_fs_forward_host_async / host_wait: five calls with different inputs, two in flight at any time; every
    result equals the blocking host call's (bitwise: same kernels, same plan)."""
    sd, src, lens, S, cfg, g = load_case("ragged_S6")
    m = make_model(sd)
    nat = m.native()
    xs = [(torch.cat(src) * (1.0 + 0.1 * i)).contiguous().pin_memory() for i in range(5)]
    want = [nat.forward_host(x, lens, S)[0] for x in xs]
    outs = [torch.empty(len(lens), max(lens), S).pin_memory() for _ in range(5)]
    tickets = []
    for i, x in enumerate(xs):
        tickets.append(nat.forward_host_async(x, lens, S, outs[i]))
        if i >= 1:
            nat.host_wait(tickets[i - 1])
            assert torch.equal(outs[i - 1], want[i - 1])
    nat.host_wait(tickets[-1])
    assert torch.equal(outs[-1], want[-1])
    nat.host_wait(tickets[0])                      # waiting on an old ticket is a no-op
    with pytest.raises(Exception):
        nat.host_wait(10 ** 6)


@pytest.mark.parametrize("chunks", [1, 2, 4])
def test_host_entry_point_chunked_copy_compute_overlap(chunks):
    """forward_host splits the batch into chunks of whole sequences (copy of chunk i+1 under the kernels of chunk i);
    the chunks share one padded T, so logits / emb / attractors land exactly where the unchunked call puts them."""
    sd, src, lens, S, cfg, g = load_case("ragged_S6")
    m = make_model(sd)
    src4 = (list(src) * 4)[:4]                      # 4 ragged sequences, longest not in the first chunk
    src4 = [src4[1], src4[0], src4[3], src4[2]] if len(src4[0]) >= len(src4[1]) else src4
    lens4 = [len(t) for t in src4]
    x = torch.cat(src4).contiguous()
    nat = m.native()
    ref_l, ref_e, ref_a = nat.forward(x.cuda(), lens4, S, want_emb=True, want_att=True)
    nat.set_option("host_chunks", chunks)
    l, e, a = nat.forward_host(x.pin_memory(), lens4, S, want_emb=True, want_att=True)
    nat.set_option("host_chunks", 0)
    for b, n in enumerate(lens4):                   # rows beyond a sequence's length are padding
        assert (l[b, :n] - ref_l.cpu()[b, :n]).abs().max().item() < 5e-4
        assert (e[b, :n] - ref_e.cpu()[b, :n]).abs().max().item() < 5e-4
        assert (a[b, :n] - ref_a.cpu()[b, :n]).abs().max().item() < 5e-4


def test_full_size_properties_B64_T500_S6():
    """BASELINE.json configs[1] shape.  (1) batch invariance: a sequence inside a batch of 64 gives the
    same logits as alone; (2) the oracle agrees on one sequence; (3) causality: perturbing frames >= 300
    leaves logits < 300 - 9 untouched; (4) logits are cosines: |y| <= 1."""
    sd = O.random_state_dict(seed=0, trained_like=True)
    m = make_model(sd)
    src, lens = O.synthetic_features(64, 500)
    srcc = [s.cuda() for s in src]
    y = torch.stack(m.test_logits(srcc, lens, 6))
    assert y.shape == (64, 500, 6) and torch.isfinite(y).all() and y.abs().max() <= 1.0 + 1e-3
    for b in (0, 17, 63):
        alone = m.test_logits([srcc[b]], [500], 6)[0]
        # not bit-exact: the speaker-attention tile (21 frames) a frame falls into depends on its global row, and
        # the fp16 partial row sums of P are grouped by tile column -> differences at the fp16-rounding level
        assert (alone - y[b]).abs().max().item() < 5e-4
    with torch.no_grad():
        ref = O.test(sd, [src[5]], [500], 6, O.Cfg())[0][0]
    err = (y[5].cpu() - ref).abs().max().item()
    print(f"B=64 T=500 S=6: max-abs logit error vs oracle = {err:.2e}")
    assert err < TOL
    pert = [s.clone() for s in srcc]
    for s in pert:
        s[300:] += 1.0
    y2 = torch.stack(m.test_logits(pert, lens, 6))
    assert (y2[:, :291] - y[:, :291]).abs().max().item() < 1e-6
    assert (y2[:, 300:] - y[:, 300:]).abs().max().item() > 1e-3


def test_weight_update_rebuilds_native_model():
    sd = O.random_state_dict(seed=7)
    m = make_model(sd)
    src, lens = O.synthetic_features(1, 64)
    a = m.test_logits([src[0].cuda()], lens, 4)[0]
    with torch.no_grad():
        m.cnn.bias.add_(0.5)
    b = m.test_logits([src[0].cuda()], lens, 4)[0]
    assert (a - b).abs().max().item() > 1e-4
    sd2 = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.test({k: v.cpu() for k, v in sd2.items()}, src, lens, 4, O.Cfg())[0][0]
    assert (b.cpu() - ref).abs().max().item() < TOL


@pytest.mark.parametrize("ffn,spk", [(0, 0), (1, 1), (2, 1), (3, 1), (4, 1), (5, 1)])
def test_kernel_variants_agree_with_reference(ffn, spk):
    """Every selectable kernel variant (unfused / fused / fused+multicast FFN; CUDA-core / tcgen05 speaker
    attention) meets the same 1e-3 bound."""
    sd, src, lens, S, cfg, g = load_case("ragged_S6")
    m = make_model(sd)
    m.native().set_option("ffn", ffn)
    m.native().set_option("spk", spk)
    out = m.test_logits([s.cuda() for s in src], lens, max_nspks=S)
    worst = max(float(np.abs(o.cpu().numpy() - g[f"logits_{i}"]).max()) for i, o in enumerate(out))
    print(f"ffn={ffn} spk={spk}: max-abs logit error vs reference = {worst:.2e}")
    assert worst < TOL


def _make_stream(sd):
    from nnet.model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import StreamingTransformerEDADiarization
    from nnet.utils.copy_params import copy_params_from_masked_to_streaming
    masked = make_model(sd)
    stream = StreamingTransformerEDADiarization(in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                                                dropout=0.1, has_mask=True, max_seqlen=500,
                                                dec_dim_feedforward=2048).cuda().eval()
    copy_params_from_masked_to_streaming(masked, stream)
    return masked, stream


def _run_stream(stream, x, S):
    """The reference's frame loop + flush (FS-EEND/streaming_infer_dia.py:77-86).  x: (B, T, 345) on the GPU."""
    ys = []
    for t in range(x.shape[1]):
        y = stream.test(x[:, t:t + 1], max_nspks=S)
        assert (y is None) == (t < 9)
        if y is not None:
            ys.append(y)
    for _ in range(9):
        ys.append(stream.test(torch.zeros(x.shape[0], 1, 345, device="cuda"), max_nspks=S, dummy_conv_input=True))
    return torch.cat(ys, dim=1)


@pytest.mark.parametrize("rv", ["1", "0"])
def test_streaming_matches_reference_stream_golden(monkeypatch, rv):
    """rv=1: the small-row CUDA-core step (default for B * S <= 16 rows); rv=0: the 128-row tcgen05 tile kernels."""
    monkeypatch.setenv("FSEEND_STREAM_RV", rv)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stream_T60_S6.npz"))
    sd = O.random_state_dict(seed=4, trained_like=True)
    masked, stream = _make_stream(sd)
    src, _ = O.synthetic_features(1, 60)
    ys = _run_stream(stream, src[0][None].cuda(), 6)[0].cpu().numpy()
    assert ys.shape == g["stream"].shape
    err = np.abs(ys - g["stream"]).max()
    print(f"streaming T=60: max-abs logit error vs reference frame loop = {err:.2e}")
    assert err < TOL


def test_streaming_infer_dia_body_on_wide_logits():
    """The body of FS-EEND/streaming_infer_dia.py:42-97 on synthetic features: masked model -> load_state_dict with the
    Lightning 'model.' prefix stripped -> masked test -> copy_params -> frame loop -> flush -> the script's own
    acceptance check torch.allclose(sigmoid(stream), sigmoid(masked), atol=1e-4, rtol=1e-4); both also against the
    real reference's outputs for the same weights (golden)."""
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    from nnet.model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import StreamingTransformerEDADiarization
    from nnet.utils.copy_params import copy_params_from_masked_to_streaming
    sd, src, lens, S, g = load_wide_case("wide_pos_S6")
    device = torch.device("cuda:0")
    params = dict(n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1, has_mask=True, max_seqlen=500,
                  dec_dim_feedforward=2048, conv_delay=9, mask_delay=0)
    feat = src[0][:120].to(device)
    masked_model = OnlineTransformerDADiarization(n_speakers=4, in_size=345, **params).to(device)
    streaming_model = StreamingTransformerEDADiarization(in_size=345, **params).to(device)
    state_dict = {"model." + k: v for k, v in sd.items()}                     # a Lightning checkpoint's key layout
    new_state_dict = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in state_dict.items()}
    masked_model.load_state_dict(new_state_dict)
    masked_model.eval()
    with torch.no_grad():
        masked_pred, _, _ = masked_model.test([feat], [len(feat)], max_nspks=S)
        masked_logits = masked_pred[0].detach().cpu().float()
        masked_pred = torch.sigmoid(masked_pred[0][:, 1:]).detach().cpu().float()
    copy_params_from_masked_to_streaming(masked_model, streaming_model)
    preds = []
    streaming_model.eval()
    with torch.no_grad():
        for t in range(len(feat)):
            pred_t = streaming_model.test(feat[t:t + 1].unsqueeze(0), max_nspks=S)
            if pred_t is not None:
                preds.append(pred_t)
        for _ in range(params["conv_delay"]):
            dummy_feat = torch.zeros(1, 1, feat.shape[-1], device=feat.device)
            pred_t = streaming_model.test(dummy_feat, max_nspks=S, dummy_conv_input=True)
            if pred_t is not None:
                preds.append(pred_t)
    preds = torch.cat(preds, dim=1)
    stream_logits = preds[0].detach().cpu().float()
    pred = torch.sigmoid(preds[0][:, 1:]).detach().cpu().float()
    print(f"stream vs masked (sigmoid space) max diff {float((pred - masked_pred).abs().max()):.2e}; vs reference: "
          f"stream {float((stream_logits - torch.from_numpy(g['stream'])).abs().max()):.2e}, "
          f"masked {float((masked_logits - torch.from_numpy(g['stream_masked'])).abs().max()):.2e} (logits)")
    assert torch.allclose(pred, masked_pred, atol=1e-4, rtol=1e-4)            # streaming_infer_dia.py:97
    assert (stream_logits - torch.from_numpy(g["stream"])).abs().max().item() < TOL
    assert (masked_logits - torch.from_numpy(g["stream_masked"])).abs().max().item() < TOL
    assert torch.allclose(pred, torch.sigmoid(torch.from_numpy(g["stream"])[:, 1:]), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("rv", ["1", "0"])
def test_streaming_equals_batch_path_two_recordings_and_cache_growth(monkeypatch, rv):
    monkeypatch.setenv("FSEEND_STREAM_RV", rv)
    _streaming_equals_batch()


def _streaming_equals_batch():
    """The reference's own invariant (streaming_infer_dia.py:97): frame-by-frame == batch.  Two recordings in
    parallel, 1100 frames (> the initial 1024-frame cache capacity, so the caches are re-allocated mid-stream)."""
    sd = O.random_state_dict(seed=9, trained_like=True)
    masked, stream = _make_stream(sd)
    T = 1100
    src, lens = O.synthetic_features(2, T)
    x = torch.stack(src).cuda()
    ys = _run_stream(stream, x, 4)
    batch = torch.stack(masked.test_logits([s.cuda() for s in src], lens, 4))
    err = (ys - batch).abs().max().item()
    print(f"streaming vs batch, B=2 T={T}: max-abs difference = {err:.2e}")
    assert ys.shape == batch.shape and err < TOL
    stream.reset()
    again = _run_stream(stream, x[:, :40], 4)
    assert (again[:, :31] - ys[:, :31]).abs().max().item() < 1e-6    # deterministic restart after reset()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_second_gpu_while_first_is_current():
    """A model that lives on cuda:1 while cuda:0 is the current device (advisor finding, round 1): the native model is
    built on the parameters' device, kernels launch there, per-device launch setup (dynamic-smem opt-in, work-queue
    counters) is repeated for the new device; results equal the cuda:0 run bit for bit."""
    sd, src, lens, S, cfg, g = load_case("ragged_S6")
    torch.cuda.set_device(0)
    m0 = make_model(sd)
    out0 = m0.test_logits([s.cuda(0) for s in src], lens, S)
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    m1 = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4,
                                        dec_n_layers=2, dropout=0.1, has_mask=True, max_seqlen=500,
                                        dec_dim_feedforward=2048)
    m1.load_state_dict(sd, strict=True)
    m1 = m1.to("cuda:1").eval()
    assert torch.cuda.current_device() == 0
    out1 = m1.test_logits([s.to("cuda:1") for s in src], lens, S)
    assert all(o.device.index == 1 for o in out1)
    for a, b in zip(out0, out1):
        assert torch.equal(a.cpu(), b.cpu())
