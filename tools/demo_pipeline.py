"""End-to-end demo of the widened path on synthetic input: 23-dim log-mel frames (10 ms) -> splice + subsample (device)
-> FS-EEND forward (device) -> sigmoid -> threshold + median filter (device) -> RTTM lines, and the same recording frame
by frame through the streaming model.  Mirrors the flow of the reference's streaming_infer_dia.py:31-104.

    python tools/demo_pipeline.py [n_mel_frames]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch  # noqa: E402

from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization  # noqa: E402
from nnet.model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import StreamingTransformerEDADiarization  # noqa: E402
from nnet.utils.copy_params import copy_params_from_masked_to_streaming  # noqa: E402
from fseend_b200.rttm import make_rttm  # noqa: E402
from fseend_b200 import feature  # noqa: E402



def main():
    n_mel = int(sys.argv[1]) if len(sys.argv) > 1 else 6000          # one minute of 10-ms frames
    torch.manual_seed(0)
    mel = torch.randn(n_mel, 23)
    feat = feature.splice_subsample(mel, context_size=7, subsampling=10)          # (T, 345) on the GPU
    kw = dict(in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1, has_mask=True,
              max_seqlen=500, dec_dim_feedforward=2048)
    masked = OnlineTransformerDADiarization(n_speakers=4, **kw).cuda().eval()
    with torch.no_grad():
        logits, _, _ = masked.test([feat], [feat.shape[0]], max_nspks=6)
    pred = torch.sigmoid(logits[0][:, 1:])
    rttm = make_rttm(rec="demo", pred=pred, frame_shift=80, subsampling=10, sampling_rate=8000)
    n_lines = sum(len(v) for v in rttm.values())
    print(f"batch: {feat.shape[0]} frames -> logits {tuple(logits[0].shape)}, {n_lines} RTTM lines; first:",
          next(iter(rttm.values()))[0] if n_lines else "-")

    stream = StreamingTransformerEDADiarization(**kw).cuda().eval()
    copy_params_from_masked_to_streaming(masked, stream)
    preds = []
    with torch.no_grad():
        for t in range(feat.shape[0]):
            y = stream.test(feat[t:t + 1].unsqueeze(0), max_nspks=6)
            if y is not None:
                preds.append(y)
        for _ in range(stream.delay if hasattr(stream, "delay") else 9):
            y = stream.test(torch.zeros(1, 1, 345, device="cuda"), max_nspks=6, dummy_conv_input=True)
            if y is not None:
                preds.append(y)
    ys = torch.cat(preds, dim=1)[0]
    diff = (ys - logits[0]).abs().max().item()
    print(f"streaming: {ys.shape[0]} frames, max |stream - batch| = {diff:.2e}")
    assert ys.shape == logits[0].shape and diff < 1e-3


if __name__ == "__main__":
    main()
