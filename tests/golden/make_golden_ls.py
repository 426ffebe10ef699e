"""Golden vectors for the LS-EEND path, produced by the REAL reference (/root/reference/LS-EEND, read-only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ls.py

Separate from make_golden.py because both reference trees use the top-level package name ``nnet``.
Weights: oracle.ls_eend_oracle.random_state_dict(seed) loaded strictly into the reference model (pins the
state_dict ABI); inputs: synthetic_features (seed 777).  Batch path = model.test(); streaming path = the frame loop
of LS-EEND/streaming_infer_dia.py:52-97 (enc.forward_one_step / StreamingConv1d / dec.forward_one_step + flush).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/LS-EEND")

from oracle import fs_eend_oracle as FO  # noqa: E402
from oracle import ls_eend_oracle as O  # noqa: E402

from nnet.model.onl_conformer_retention_enc_1dcnn_tfm_retention_enc_linear_non_autoreg_pos_enc_l2norm_emb_loss_mask import (  # noqa: E402
    OnlineConformerRetentionDADiarization, StreamingConv1d)

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (weight seed, trained_like, lens, max_nspks)
    "ls_T1000_ragged_S6": (0, True, [1000, 730], 6),      # two retention chunks, ragged batch
    "ls_T300_S4": (1, True, [300], 4),                    # single (zero-padded) chunk
    "ls_T1200_S10": (2, True, [1200], 10),                # three chunks, 8-speaker slot count
}


def build_ref(sd):
    m = OnlineConformerRetentionDADiarization(
        n_speakers=8, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1,
        max_seqlen=1000, recurrent_chunk_size=500, feed_forward_expansion_factor=4, dec_dim_feedforward=2048,
        conv_expansion_factor=2, conv_kernel_size=16, half_step_residual=True, conv_delay=9)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def streaming_predict(model, feat, max_nspks):
    """LS-EEND/streaming_infer_dia.py:52-97, B = 1."""
    n_enc, n_dec = len(model.enc.encoder.layers), len(model.dec.layers)
    ret_states = [dict() for _ in range(n_enc)]
    conv_caches = [torch.zeros(1, model.n_units, model.enc.encoder._conv_kernel_size - 1) for _ in range(n_enc)]
    dec_states = [dict() for _ in range(n_dec)]
    cnn = StreamingConv1d(model.n_units, model.n_units, kernel_size=2 * model.delay + 1).eval()
    cnn.conv.load_state_dict(model.cnn.state_dict())
    preds, dec_t = [], 0

    def step(emb_t, dec_t):
        e = cnn(emb_t.transpose(1, 2))
        if e is None:
            return None, dec_t
        e = e.transpose(1, 2)
        e = e / torch.norm(e, dim=-1, keepdim=True)
        a = model.dec.forward_one_step(e, dec_t, max_nspks, dec_states)
        a = a / torch.norm(a, dim=-1, keepdim=True)
        return torch.matmul(e.unsqueeze(-2), a.transpose(-1, -2)).squeeze(-2), dec_t + 1

    for t in range(feat.shape[0]):
        emb_t = model.enc.forward_one_step(feat[t:t + 1].unsqueeze(0), t, ret_states, conv_caches)
        y, dec_t = step(emb_t, dec_t)
        if y is not None:
            preds.append(y)
    for _ in range(model.delay):
        y, dec_t = step(torch.zeros(1, 1, model.n_units), dec_t)
        if y is not None:
            preds.append(y)
    return torch.cat(preds, dim=1).squeeze(0)


def forward_loss_golden():
    """Training-signature forward (eval arithmetic) with labels: the length-masked emb-consistency loss (LS:model:92-113)."""
    sd = O.random_state_dict(seed=0, trained_like=True)
    ref = build_ref(sd)
    lens = [700, 433]
    S = 6
    src, lens = FO.synthetic_features(len(lens), max(lens), lens=lens)
    g = torch.Generator().manual_seed(123)
    tgt = [(torch.rand(l, n, generator=g) > 0.6).float() for l, n in zip(lens, (S, S - 2))]
    with torch.no_grad():
        out, loss, emb, att = ref(src, tgt, lens)
    rec = {"emb_consis_loss": np.array(loss.item(), dtype=np.float64), "fwd_logits_0": out[0].numpy(),
           "fwd_logits_1": out[1].numpy(), "att_shape_1": np.array(att[1].shape)}
    np.savez_compressed(os.path.join(HERE, "ls_forward_loss_S6.npz"), **rec)
    print("ls_forward_loss_S6: loss", loss.item(), [tuple(o.shape) for o in out], tuple(att[1].shape))


BIG_CASES = {
    # BASELINE.json configs[2] shape: B=16 recordings x T=2000 frames, 8 speakers (+2 slots); one recording shorter
    "ls_B16_T2000_S10": (4, True, [2000] * 15 + [1711], 10),
}


def big_goldens():
    """BASELINE-shape batch golden (logits only, 1.3 MB) and a long one-step golden (T = 2000 frames through the real
    streaming_predict loop: 4 retention chunks' worth of recurrent state, S = 10)."""
    for name, (wseed, trained, lens, S) in BIG_CASES.items():
        sd = O.random_state_dict(seed=wseed, trained_like=trained)
        ref = build_ref(sd)
        src, lens = FO.synthetic_features(len(lens), max(lens), lens=lens)
        with torch.no_grad():
            out, _, _ = ref.test(src, lens, max_nspks=S)
        rec = {f"logits_{i}": o.numpy() for i, o in enumerate(out)}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, [tuple(o.shape) for o in out][:2], "max|logit|", max(o.abs().max().item() for o in out))
    sd = O.random_state_dict(seed=5, trained_like=True)
    ref = build_ref(sd)
    src, lens = FO.synthetic_features(1, 2000)
    with torch.no_grad():
        ys = streaming_predict(ref, src[0], 10)
    np.savez_compressed(os.path.join(HERE, "ls_stream_T2000_S10.npz"), stream=ys.numpy())
    print("ls_stream_T2000_S10", tuple(ys.shape), "max|logit|", ys.abs().max().item())


def main():
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "forward_loss":
        forward_loss_golden()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        big_goldens()
        return
    forward_loss_golden()
    for name, (wseed, trained, lens, S) in CASES.items():
        sd = O.random_state_dict(seed=wseed, trained_like=trained)
        ref = build_ref(sd)
        src, lens = FO.synthetic_features(len(lens), max(lens), lens=lens)
        with torch.no_grad():
            out, emb, att = ref.test(src, lens, max_nspks=S)
        rec = {f"logits_{i}": o.numpy() for i, o in enumerate(out)}
        rec.update({f"emb_{i}": e.numpy()[:: max(1, len(e) // 8)] for i, e in enumerate(emb)})
        rec["emb_stride"] = np.array([max(1, len(e) // 8) for e in emb])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, [tuple(o.shape) for o in out], "max|logit|", max(o.abs().max().item() for o in out))

    sd = O.random_state_dict(seed=3, trained_like=True)
    ref = build_ref(sd)
    src, lens = FO.synthetic_features(1, 48)
    with torch.no_grad():
        ys = streaming_predict(ref, src[0], 4)
        batch = ref.test(src, lens, max_nspks=4)[0][0]
    print("LS stream vs batch max diff", (ys - batch).abs().max().item())
    np.savez_compressed(os.path.join(HERE, "ls_stream_T48_S4.npz"), stream=ys.numpy(), batch=batch.numpy())

    with open(os.path.join(HERE, "ls_state_dict_abi.txt"), "w") as f:
        for k, v in ref.state_dict().items():
            f.write(f"{k} {tuple(v.shape)} {str(v.dtype).replace('torch.', '')}\n")


if __name__ == "__main__":
    main()
