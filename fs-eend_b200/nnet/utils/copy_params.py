"""copy_params_from_masked_to_streaming (API mirror of the reference's FS-EEND/nnet/utils/copy_params.py:59-62):
copies every parameter/buffer of the masked (batch) model into the frame-by-frame model, using the name map that is
the two models' ABI."""
import torch

from ..model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import streaming_to_masked_key


@torch.no_grad()
def copy_params_from_masked_to_streaming(masked_fs_eend, streaming_fs_eend):
    src = masked_fs_eend.state_dict()
    dst = streaming_fs_eend.state_dict()
    for k, v in dst.items():
        v.copy_(src[streaming_to_masked_key(k)])
    if hasattr(streaming_fs_eend, "invalidate_native"):
        streaming_fs_eend.invalidate_native()      # weights changed: the cached native model is stale by definition
    if hasattr(streaming_fs_eend, "reset"):
        streaming_fs_eend.reset()
