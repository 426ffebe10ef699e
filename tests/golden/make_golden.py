"""Generate golden vectors by running the REAL reference (imported read-only from /root/reference).

Run once in the authoring container (the reference tree does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Weights come from ``oracle.fs_eend_oracle.random_state_dict(seed)`` (deterministic CPU generator,
loaded into the reference modules by name, strict=True — this also pins the state_dict ABI),
inputs from ``synthetic_features`` (seed 777).  Outputs are stored as float32
``.npz`` files small enough to commit.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = "/root/reference/FS-EEND"
sys.path.insert(0, REF)

from oracle import fs_eend_oracle as O  # noqa: E402

from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization  # noqa: E402
from nnet.model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import StreamingTransformerEDADiarization  # noqa: E402
from nnet.utils.copy_params import copy_params_from_masked_to_streaming  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def build_ref(sd, mask_delay=0, n_speakers=4):
    m = OnlineTransformerDADiarization(
        n_speakers=n_speakers, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
        dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048, mask_delay=mask_delay)
    m.load_state_dict(sd, strict=True)
    return m.eval()


CASES = {
    # name: (weight seed, trained_like, lens, max_nspks, mask_delay)
    "c1_T500_S4": (0, True, [500], 4, 0),                 # BASELINE configs[0]
    "c1_T500_S6": (0, True, [500], 6, 0),                 # north-star shape at B=1
    "ragged_S6": (1, True, [200, 137, 64, 19], 6, 0),     # ragged batch
    "default_init_S4": (2, False, [96, 96], 4, 0),
    "maskdelay2_S5": (3, True, [150, 90], 5, 2),
}


WIDE_CASES = {
    # name: (weight seed, alpha, lens, max_nspks): logits spanning most of the cosine range (oracle.wide_state_dict)
    "wide_pos_S6": (0, 1.5, [300, 211], 6),
    "wide_neg_S6": (1, -1.5, [300, 211], 6),
}


def wide_goldens():
    """Wide-dynamic-range weight sets: masked-model logits, plus (for the first) the streaming model's frame loop and the
    reference's own sigmoid-space self-check (streaming_infer_dia.py:97: allclose(stream, masked, atol=rtol=1e-4))."""
    for name, (wseed, alpha, lens, S) in WIDE_CASES.items():
        sd = O.wide_state_dict(seed=wseed, alpha=alpha, S=S)
        ref = build_ref(sd)
        src, lens = O.synthetic_features(len(lens), max(lens), lens=lens)
        with torch.no_grad():
            out, _, _ = ref.test(src, lens, max_nspks=S)
        rec = {f"logits_{i}": o.numpy() for i, o in enumerate(out)}
        if alpha > 0:
            stream = StreamingTransformerEDADiarization(
                in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1, has_mask=True,
                max_seqlen=500, dec_dim_feedforward=2048).eval()
            copy_params_from_masked_to_streaming(ref, stream)
            T = 120
            with torch.no_grad():
                ys = []
                for t in range(T):
                    y = stream.test(src[0][None, t:t + 1], max_nspks=S)
                    if y is not None:
                        ys.append(y)
                for _ in range(9):
                    ys.append(stream.test(torch.zeros(1, 1, 345), max_nspks=S, dummy_conv_input=True))
                ys = torch.cat(ys, dim=1)[0]
                masked = ref.test([src[0][:T]], [T], max_nspks=S)[0][0]
            ok = torch.allclose(torch.sigmoid(ys[:, 1:]), torch.sigmoid(masked[:, 1:]), atol=1e-4, rtol=1e-4)
            print(name, "reference self-check (sigmoid space, 1e-4):", ok, (ys - masked).abs().max().item())
            rec["stream"] = ys.numpy()
            rec["stream_masked"] = masked.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        y = torch.cat([o.flatten() for o in out])
        print(name, "logit range", y.min().item(), y.max().item(), "std", y.std().item())


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "wide":
        wide_goldens()
        return
    for name, (wseed, trained, lens, S, md) in CASES.items():
        sd = O.random_state_dict(seed=wseed, trained_like=trained)
        ref = build_ref(sd, mask_delay=md)
        src, lens = O.synthetic_features(len(lens), max(lens), lens=lens)
        with torch.no_grad():
            out, emb, att = ref.test(src, lens, max_nspks=S)
        rec = {f"logits_{i}": o.numpy() for i, o in enumerate(out)}
        rec.update({f"emb_{i}": e.numpy()[:: max(1, len(e) // 8)] for i, e in enumerate(emb)})   # subsampled rows
        rec["emb_stride"] = np.array([max(1, len(e) // 8) for e in emb])
        # training-mode signature (eval arithmetic): forward with labels -> emb consistency loss
        g = torch.Generator().manual_seed(123)
        tgt = [(torch.rand(l, S, generator=g) > 0.6).float() for l in lens]
        with torch.no_grad():
            fout, loss, _, _ = ref(src, tgt, lens)
        rec["emb_consis_loss"] = np.array(loss.item(), dtype=np.float64)
        rec["fwd_logits_0"] = fout[0].numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, "logits[0] shape", out[0].shape, "loss", loss.item())

    # streaming path: reference frame loop (FS-EEND/streaming_infer_dia.py:77-86) on T=60
    sd = O.random_state_dict(seed=4, trained_like=True)
    ref = build_ref(sd)
    stream = StreamingTransformerEDADiarization(
        in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1, has_mask=True,
        max_seqlen=500, dec_dim_feedforward=2048).eval()
    copy_params_from_masked_to_streaming(ref, stream)
    src, lens = O.synthetic_features(1, 60)
    with torch.no_grad():
        batch = ref.test(src, lens, max_nspks=6)[0][0]
        ys = []
        for t in range(60):
            y = stream.test(src[0][None, t:t + 1], max_nspks=6)
            if y is not None:
                ys.append(y)
        for _ in range(9):
            ys.append(stream.test(torch.zeros(1, 1, 345), max_nspks=6, dummy_conv_input=True))
        ys = torch.cat(ys, dim=1)[0]
    print("stream vs batch max diff", (ys - batch).abs().max().item())
    np.savez_compressed(os.path.join(HERE, "stream_T60_S6.npz"), stream=ys.numpy(), batch=batch.numpy())

    with open(os.path.join(HERE, "fs_stream_state_dict_abi.txt"), "w") as f:
        for k, v in stream.state_dict().items():
            f.write(f"{k} {tuple(v.shape)} {str(v.dtype).replace('torch.', '')}\n")

    # state_dict ABI: names and shapes of the reference model
    with open(os.path.join(HERE, "fs_state_dict_abi.txt"), "w") as f:
        for k, v in build_ref(O.random_state_dict(0)).state_dict().items():
            f.write(f"{k} {tuple(v.shape)} {str(v.dtype).replace('torch.', '')}\n")


if __name__ == "__main__":
    main()
